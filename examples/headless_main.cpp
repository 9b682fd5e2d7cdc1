// headless_main.cpp -- runCuda() of the reference (main.cpp:146-185) on top of restir_shim.hpp, without GL:
//   headless <scene.txt> <frames> <out.ppm> [reuse 0..3]
// Build: g++ -std=c++17 -Iinclude examples/headless_main.cpp -Lrestir_b200 -lrestir_b200 -Wl,-rpath,$PWD/restir_b200 -o headless
#include <cmath>
#include <cstring>
#include <vector>

#include "restir_shim.hpp"

using namespace restir_shim;

int main(int argc, char** argv) {
    if (argc < 4) { std::printf("Usage: %s SCENEFILE.txt FRAMES OUT.ppm [reuse]\n", argv[0]); return 1; }
    Scene scene(argv[1]);
    const int frames = std::atoi(argv[2]);
    if (argc > 4) Settings::reservoirReuse = std::atoi(argv[4]);
    Settings::animateCamera = true;
    Camera& cam = scene.camera;
    const int width = cam.resolution[0], height = cam.resolution[1];
    GBuffer gBuffer;
    gBuffer.create(scene, width, height);
    std::vector<unsigned char> pixels((size_t)width * height * 4);
    int iteration = 0;
    for (int k = 0; k < frames; k++) {                              // one runCuda() per iteration
        float orig[3];
        std::memcpy(orig, cam.position, sizeof orig);
        if (Settings::animateCamera) {                              // main.cpp:149-153 with the clock t_k = k * speed / 60
            float t = float(k) / 60.f * Settings::animateSpeed;
            cam.position[0] = orig[0] + std::cos(t) * Settings::animateRadius;
            cam.position[2] = orig[2] + std::sin(t) * Settings::animateRadius;
        }
        if (!Settings::accumulate) iteration = 0;                   // main.cpp:155-162
        check(rstr_camera_update(&cam));
        gBuffer.render(cam);
        if (Settings::useReservoir) ReSTIRDirect(gBuffer, cam, iteration);
        else pathTraceDirect(gBuffer, cam, iteration);
        copyImageToHost(gBuffer, pixels.data(), pixels.size(), Settings::toneMapping);
        iteration++;
        gBuffer.update(cam);
        std::memcpy(cam.position, orig, sizeof orig);               // main.cpp:184
    }
    FILE* f = std::fopen(argv[3], "wb");                            // saveImage (main.cpp:105-144) mirrors x; PPM keeps it simple
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) std::fwrite(&pixels[4 * ((size_t)y * width + (width - 1 - x))], 1, 3, f);
    std::fclose(f);
    unsigned long long mean = 0;
    for (size_t i = 0; i < pixels.size(); i += 4) mean += pixels[i] + pixels[i + 1] + pixels[i + 2];
    std::printf("{\"frames\": %d, \"width\": %d, \"height\": %d, \"mean_ldr\": %.3f}\n", frames, width, height, double(mean) / (3.0 * width * height));
    gBuffer.destroy();
    scene.clear();
    return 0;
}
