// camera_dev.h -- the camera's ray generation on the device, shared by kernels.cu and denoise.cu
#pragma once

#include "device_types.h"
#include "vecmath.h"

namespace rs {

// Camera::sample (sceneStructs.h:69-86); gbuffer.cu:11-23 is the same expression with r = (.5, .5)
RS_D void cameraRay(const CamDev& c, int x, int y, float rx, float ry, f3& o, f3& d) {
    float sx = (float)x * c.pixelSizeX, sy = (float)y * c.pixelSizeY;
    float ux = sx + c.pixelSizeX * rx, uy = sy + c.pixelSizeY * ry;
    ux = 1.f - ux * 2.f; uy = 1.f - uy * 2.f;
    f3 pf = mk3(ux * c.aspect * c.tanFovY, uy * 1.f * c.tanFovY, 1.f) * c.focalDist;
    f3 dir = pf - mk3(0.f);
    f3 right = mk3(c.right[0], c.right[1], c.right[2]), up = mk3(c.up[0], c.up[1], c.up[2]), view = mk3(c.view[0], c.view[1], c.view[2]);
    f3 w = mk3(right.x * dir.x + up.x * dir.y + view.x * dir.z,
               right.y * dir.x + up.y * dir.y + view.y * dir.z,
               right.z * dir.x + up.z * dir.y + view.z * dir.z);
    d = normalize(w);
    o = mk3(c.position[0], c.position[1], c.position[2]) + right * 0.f + up * 0.f;
}
// the origin cameraRay gives every ray (the same expression: warp-uniform, the packet walk keeps it out of per-thread registers)
RS_D f3 cameraOrigin(const CamDev& c) {
    f3 right = mk3(c.right[0], c.right[1], c.right[2]), up = mk3(c.up[0], c.up[1], c.up[2]);
    return mk3(c.position[0], c.position[1], c.position[2]) + right * 0.f + up * 0.f;
}

// Camera::getPosition (sceneStructs.h:48-64): the point at distance `dist` along the pixel-centre ray -- the same expressions as
// Camera::sample with r = (.5, .5) and no lens
RS_D f3 cameraPosition(const CamDev& c, int x, int y, float dist) {
    f3 o, d;
    cameraRay(c, x, y, .5f, .5f, o, d);
    return o + d * dist;
}

}  // namespace rs
