// restir_shim.hpp -- the reference's frame interface (names and argument meaning of gbuffer.h:24-27, restir.h:128-132,
// pathtrace.h:8-16, scene.h:487-491) as a thin C++ layer over the C ABI of librestir_b200.so.  This is the file a
// maintainer of HummaWhite/ReSTIR adds next to main.cpp; it needs no CUDA headers and no nvcc.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <string>

#include "restir_b200.h"

namespace restir_shim {

// common.h:36-43, 47-67: the run-time toggles the UI edits
struct ReservoirReuse { enum { None = 0, Temporal = 1, Spatial = 2, Spatiotemporal = 3 }; };
struct Settings {
    static inline bool useReservoir = true;
    static inline int reservoirReuse = ReservoirReuse::Temporal;     // common.cpp:14
    static inline bool accumulate = false;
    static inline int toneMapping = 2;                               // ACES, common.cpp:4
    static inline bool animateCamera = false;
    static inline float animateRadius = 1.f, animateSpeed = 2.7f;
};
struct State { static inline int looper = 0; };

using Camera = RstrCamera;   // same 196-byte POD as sceneStructs.h:22-126

inline void check(int rc) {  // cudaUtil.h:13-31: print and exit
    if (rc != RSTR_OK) { std::fprintf(stderr, "restir_b200 error %d: %s\n", rc, rstr_last_error()); std::exit(EXIT_FAILURE); }
}

class Scene {                // scene.h:483-531
public:
    explicit Scene(const std::string& filename) { check(rstr_init(0)); check(rstr_scene_load_file(filename.c_str(), &handle, &camera)); }
    void buildDevData() {}   // done by the constructor (host build) and by the first frame (upload)
    void clear() { rstr_scene_destroy(handle); handle = nullptr; }
    Camera camera;
    RstrScene* handle = nullptr;
};

class GBuffer {              // gbuffer.h:15-58
public:
    void create(Scene& scene, int width, int height) { check(rstr_frame_create(scene.handle, width, height, &frame)); }
    void destroy() { rstr_frame_destroy(frame); frame = nullptr; }
    void render(const Camera& cam) { check(rstr_gbuffer_render(frame, &cam)); }
    void update(const Camera& cam) { check(rstr_gbuffer_update(frame, &cam)); }
    RstrFrame* frame = nullptr;
};

inline void ReSTIRReset(GBuffer& g) { check(rstr_frame_reset(g.frame)); }                      // restir.cu:516
inline void ReSTIRDirect(GBuffer& g, const Camera& cam, int iter) {                             // restir.cu:418-446
    RstrParams p;
    rstr_params_default(&p);
    p.reuse = Settings::reservoirReuse;
    check(rstr_restir_direct(g.frame, &cam, &p, State::looper, iter));
    State::looper++;
}
inline void pathTraceDirect(GBuffer& g, const Camera& cam, int iter) {                          // pathtrace.cu:457-473
    check(rstr_pathtrace_direct(g.frame, &cam, State::looper, iter));
    State::looper++;
}
inline void copyImageToHost(GBuffer& g, void* rgba8, size_t bytes, int toneMapping) {           // pathtrace.cu:108-113
    check(rstr_tonemap(g.frame, toneMapping, 1.f));
    check(rstr_frame_read(g.frame, RSTR_BUF_LDR, rgba8, bytes));
}

}  // namespace restir_shim
