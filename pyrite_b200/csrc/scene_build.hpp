// Host scene builder: project IR -> the flat arrays the sm_100a kernels read (device_types.h).
// This is the product's equivalent of `parse_project` (pyrite/src/main.rs:111-134):
// Camera::from_project (cameras.rs:30-55), Renderer::from_project (renderer/mod.rs:31-75),
// World::from_project (world.rs:39-271) with Material::from_project (materials/mod.rs:33-228),
// ProgramCompiler::compile (program/compiler.rs:48-586) and Bvh::new (spatial/bvh.rs:13-155).
// It runs once per project on the host; nothing in it is on the per-sample path.
#pragma once
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "bvh_build.hpp"
#include "device_types.h"
#include "project_ir.hpp"

namespace pyr {

// std::vector whose resize() does not zero-fill trivial records: the per-primitive arrays of a big mesh (hundreds of MB) are written
// in full by several threads right after they are sized, and those threads - not one zero-filling thread - should touch the pages first.
template <class T> struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    using std::allocator<T>::allocator;
    // (records of plain floats / integers with default member initialisers are not "trivial" for the language, but nothing reads an
    // entry before the filling loop has assigned it: such entries are not constructed at all)
    template <class U> void construct(U* p) noexcept {
        if constexpr (!(std::is_trivially_copyable<U>::value && std::is_trivially_destructible<U>::value)) ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template <class T> using RawVector = std::vector<T, DefaultInitAllocator<T>>;

struct BakedScene {
    RawVector<Node4> nodes;
    RawVector<Prim> prims;            // leaf pre-order ("rank")
    RawVector<TriShade> tri_shade;    // by rank
    RawVector<TriFrames> tri_frames;  // by rank; empty unless some material has a normal map
    std::vector<PlaneRec> planes;
    std::vector<MarchedRec> marched;
    std::vector<MaterialRec> materials;
    std::vector<ComponentRec> components;
    std::vector<ProgramRec> programs;
    std::vector<Instr> code;
    std::vector<SpectrumRec> spectra;
    std::vector<float> spectrum_data;
    std::vector<TextureRec> textures;
    std::vector<float> texels;
    std::vector<LampRec> lamps;
    std::vector<TileRec> tiles;
    std::vector<float> burns, xyz, d65;
    std::vector<uint32_t> rank_of_object;  // object id -> rank
    uint32_t n_objects = 0;
    uint32_t bvh_depth = 0;
    // header part of SceneView (pointers are filled in after upload)
    SceneView view{};
};

// Throws ir::BuildError on a malformed project (missing mesh material, vector used as number...).
// `bvh_builder`: who builds the BVH of two items or more (nullptr: the depth-first host builder in scene_build.cpp; the C ABI
// passes the GPU build of bvh_build.cu - both make the same tree, tests/test_host_logic.py and tests/test_gpu_parity.py).
BakedScene build_scene(const ir::Document& doc, const BvhBuildFn* bvh_builder = nullptr);

// renderer/algorithm.rs:152-188 `make_tiles` + cameras.rs:57-68 `to_view_area` (row-major order;
// the reference's centre-out sort only changes scheduling order).
std::vector<TileRec> make_tiles(uint32_t width, uint32_t height, uint32_t tile_size);

}  // namespace pyr
