// The BVH of a scene as the builders hand it to scene_build.cpp: the binary tree of Bvh::new (spatial/bvh.rs:13-155) with its
// leaves in the reference's flattened pre-order ("ranks") and its interior nodes numbered in the same pre-order.
#pragma once
#include <stdint.h>

#include <functional>
#include <memory>
#include <vector>

namespace pyr {

struct BvhInterior {
    float box[2][6];   // child 0 / 1: lo xyz, hi xyz
    int32_t child[2];  // >= 0: interior index; < 0: leaf, rank = ~child.  child 0 is the one the reference's walk visits first
};
static_assert(sizeof(BvhInterior) == 56, "BvhInterior is 14 words");

struct BvhTree {
    std::vector<uint32_t> order;         // item index per rank
    std::unique_ptr<BvhInterior[]> interiors;  // n_interiors records, not zero-filled (49 MB for config C2's mesh); interiors[0] is the root
    size_t n_interiors = 0;
    void allocate_interiors(size_t count) { interiors.reset(new BvhInterior[count]); n_interiors = count; }
    int32_t root = 0;                    // child code of the root
    int max_depth = 0;                   // depth of the deepest leaf (root = 0)
};

// boxes6: n x (lo xyz, hi xyz), n >= 2; root_hull12: lo, hi of the union of the boxes, lo, hi of their centres.
using BvhBuildFn = std::function<void(const float* boxes6, size_t n, const float* root_hull12, BvhTree& out)>;

}  // namespace pyr
