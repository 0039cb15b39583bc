// Bidirectional integrator (renderer/bidirectional.rs) - device-side stage logic.
#pragma once
#include "shading.cuh"

namespace pyr {

// One stored vertex of the lamp subpath (tracer.rs:157-167 `Bounce`, without direct light).
struct LightVertex {
    float position[3];
    uint32_t type;        // 0 diffuse, 1 specular, 2 emission
    float normal[3];
    int32_t color_program;
    float incident[3];
    float probability;
    float out[3];         // BounceType::Diffuse(_, out)
    uint32_t dispersed;
    float tex[2];
    uint32_t pad[2];
};

struct BidirOut {
    uint32_t alive, n_rays;
    Ray rays[1 + MAX_LIGHT_PATH];
};
PYR_HD void generate_bidirectional(const SceneView&, uint64_t, uint32_t, uint64_t, PathState&, LightVertex*, BidirOut& bo) { bo.alive = 0; bo.n_rays = 0; }
template <class Add>
PYR_HD void shade_bidirectional(const SceneView&, PathState&, LightVertex*, const Ray*, const Hit*, BidirOut& bo, Add&, PathCounters&) { bo.alive = 0; bo.n_rays = 0; }

}  // namespace pyr
