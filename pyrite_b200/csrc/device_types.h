// Flat, pointer-free scene and pipeline records shared by the host scene builder (C++) and the
// sm_100a kernels (CUDA).  Everything the kernels read lives in HBM as arrays of these PODs;
// DESIGN.md §4 describes the layout and the per-unit byte counts.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PYR_HD __host__ __device__ __forceinline__
#define PYR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PYR_HD inline
#define PYR_HD_NOINLINE inline
#endif

namespace pyr {

constexpr float DIST_EPSILON = 0.0001f;  // reference: pyrite/src/math.rs:4
constexpr int MAX_SPECTRUM_SAMPLES = 16;
constexpr int MAX_LIGHT_SAMPLES = 8;
constexpr int MAX_RAY_MARCHED = 32;  // ray-marched shapes: candidate bitmask per ray
constexpr int BVH_STACK = 64;
constexpr int MAX_LIGHT_PATH = 32;   // stored lamp-subpath vertices per path sample (bidirectional)

struct alignas(16) f4 { float x, y, z, w; };

// ---- geometry --------------------------------------------------------------------------
// One BVH leaf primitive, stored in the reference BVH's leaf pre-order ("rank" order,
// spatial/bvh.rs:250-275), so that comparing indices implements World::intersect's tie rule
// (world.rs:288-296: strict `<`, earlier leaf wins).  48 B = three 16-byte loads.
//   triangle: a = (v1.xyz, e1.x)  b = (e1.y, e1.z, e2.x, e2.y)  c.x = e2.z     (shapes/mod.rs:75-119)
//   sphere:   a = (centre.xyz, radius)  b.x, b.y = texture_scale
//   marched:  a.x = index into the ray-marched table (uint bits)
//   c.y = kind (uint bits), c.z = object id (insertion order, world.rs:75,182,229), c.w = material
struct alignas(16) Prim { f4 a, b, c; };

// 4-wide BVH node: the reference's binary tree (spatial/bvh.rs:13-155) with every second level folded away.
// A node holds the boxes of up to four descendants of one binary node - its children, or its
// grandchildren where a child is itself interior - with the reference's box VALUES, so every leaf is
// still guarded by exactly its own box and its ancestors' boxes that remain.  Skipping the folded
// ancestors' tests changes nothing: a child box lies inside its parent's and the slab test is monotone.
// 128 B = eight 16-byte loads, boxes stored component-wise (lo_x[4], lo_y[4], ...).
//   child[k] >= 0: Node4 index; < 0: leaf, rank = ~child; NODE4_EMPTY: no child in this slot
struct alignas(32) Node4 {
    f4 lo_x, lo_y, lo_z, hi_x, hi_y, hi_z;
    int32_t child[4];
    uint32_t pad[4];
};
constexpr int32_t NODE4_EMPTY = 0x7fffffff;

// Per-triangle shading data in the same rank order: vertex normals, UVs and the material, 64 B - everything the shade stage needs
// of a triangle hit, so that it does not touch the 48-byte Prim record at all (42 MB less random-access data competing for L2
// on the 871k-triangle mesh).
struct alignas(16) TriShade { float n1[3], n2[3], n3[3]; float t1[2], t2[2], t3[2]; uint32_t material; };
// Tangent-space quaternions of the three vertices (read only by normal-mapped materials). 48 B.
struct TriFrames { f4 q1, q2, q3; };  // (s, x, y, z)

struct alignas(16) PlaneRec {      // shapes/mod.rs:434-439 + collision::Plane{n, d}
    float n[3], d;
    f4 from_space;     // Normal::from_space quaternion (s, x, y, z)
    float texture_scale[2];
    uint32_t material, pad;
};

struct alignas(16) MarchedRec {    // shapes/mod.rs:47-51, shapes/distance_estimators.rs, BoundingVolume :586-589
    uint32_t estimator;       // 0 mandelbulb, 1 quaternion julia
    uint32_t iterations;
    float threshold, power, slice_plane;
    uint32_t has_constant, variant;   // variant: 0 regular 1 cubic 2 bicomplex
    float mb_constant[3];
    f4 constant;              // quaternion (s, x, y, z)
    uint32_t bounds_type;     // 0 box, 1 sphere
    float ba[3], bb[3], bradius;
    uint32_t rank, object_id, material;
};

// ---- materials / programs ----------------------------------------------------------------
enum BsdfType : uint32_t { BSDF_EMISSIVE = 0, BSDF_DIFFUSE = 1, BSDF_MIRROR = 2, BSDF_REFRACTIVE = 3 };

struct alignas(16) ComponentRec {   // materials/mod.rs:230-235, 315-340
    uint32_t bsdf;
    int32_t color_program;
    int32_t probability_program;  // -1 = None
    float selection_compensation;
    float ior, env_ior, dispersion, env_dispersion;
};
struct alignas(16) MaterialRec {    // materials/mod.rs:26-30, 83-87
    uint32_t comp_offset, n_components, emissive_offset, n_emissive;
    int32_t normal_map_program;   // -1 = None
    uint32_t pad[3];
};

// Register bytecode (program/instruction.rs:19-92) re-encoded as fixed 32-byte records over one
// file of 16-byte registers (numbers live in .x; rgb = (r, g, b, alpha); vectors = xyzw).
enum Op : uint8_t {
    OP_NUMBER = 0,        // out.x = a
    OP_VECTOR,            // out = (a, b, c, d)                     instruction.rs VectorValue
    OP_RGB,               // out = (a, b, c, 1)                     RgbValue
    OP_SPECTRUM,          // out.x = spectra[resource](wavelength)  SpectrumValue
    OP_COLOR_TEXTURE,     // out = bicubic rgba                     ColorTextureValue
    OP_MONO_TEXTURE,      // out.x = bicubic luma                   MonoTextureValue
    OP_RGB_SPECTRUM,      // out.x = Burns(R[a], wavelength)        RgbSpectrumValue
    OP_FRESNEL,           // out.x = fresnel(a, b, normal, incident)
    OP_BLACKBODY,         // out.x = blackbody(wavelength, a)
    OP_NUM_TO_RGB,        // out = (n, n, n, 1)   n = R[a].x        Convert
    OP_NUM_TO_VEC,        // out = (n, n, n, n)
    OP_RGB_TO_VEC,        // out = R[a] * 2 - 1
    OP_BINARY,            // out = R[a] (binop) R[b], on `vtype` lanes
    OP_MIX,               // out = mix(R[b], R[c], clamp(a))
    OP_CLAMP              // out.x = max(min(a, c), b)
};
enum ValueType : uint8_t { VT_NUMBER = 0, VT_VECTOR = 1, VT_RGB = 2 };
enum InputBits : uint8_t { IN_WAVELENGTH = 1, IN_NORMAL = 16, IN_INCIDENT = 32, IN_TEXTURE = 64 };  // program/mod.rs:150-158

struct alignas(16) Instr {
    uint8_t op, vtype, binop, out;    // binop: 0 add 1 sub 2 mul 3 div
    uint8_t is_reg[4];                // operand a..d: 1 = register index in v[i].u, 0 = constant in v[i].f
    uint8_t deps, pad[3];             // InputBits this instruction depends on (transitively)
    uint32_t resource;                // spectrum / texture id
    union { float f; uint32_t u; } v[4];
};
static_assert(sizeof(Instr) == 32, "Instr must be 32 bytes");

constexpr int VM_REGS = 16;

struct alignas(16) ProgramRec {
    uint32_t is_constant;
    float value;
    uint32_t code_offset, n_instr;
    uint32_t out_reg;
    uint32_t reads;           // union of deps
    // the wavelength-dependent instructions again, contiguously: what a memoised re-run executes
    // (MemoizedContext, program/execution_context.rs:310-342)
    uint32_t wl_offset, wl_count;
};

struct alignas(16) SpectrumRec { uint32_t is_curve; float lo, hi; uint32_t offset, n, pad[3]; };  // curve: (x, y) pairs at offset
struct alignas(16) TextureRec { uint32_t width, height, channels, pad; uint64_t offset; uint64_t pad2; };

enum LampKind : uint32_t { LAMP_DIRECTIONAL = 0, LAMP_POINT = 1, LAMP_SHAPE = 2 };
struct alignas(16) LampRec {       // lamp.rs:11-19
    uint32_t kind;
    int32_t color_program;
    float v[3];         // direction | position
    float width;
    uint32_t rank;      // LAMP_SHAPE: primitive rank
    uint32_t pad;
};

struct CameraRec {     // cameras.rs:20-27
    float m[16];        // camera-to-world, column-major
    float inv[16];      // world-to-camera (Camera::is_visible inverts per call, cameras.rs:112)
    float view_plane, focus_distance, aperture;
    uint32_t inv_ok;
};

struct alignas(16) TileRec { float from[2], size[2]; uint32_t width, height, index, pad; };

// ---- the resolved renderer (renderer/mod.rs:18-28) ----------------------------------------
struct RendererRec {
    uint32_t algorithm, bounces, pixel_samples, light_samples, spectrum_samples, spectrum_bins, tile_size, light_bounces;
    float span_lo, span_hi;
    uint32_t width, height;
};

struct FilmRec {       // film.rs:9-18 + AspectRatio :203-224
    uint32_t width, height, bins, horizontal;
    float wavelength_start, wavelength_width, grains_per_wavelength;
    float ar_size, ar_ratio;
};

struct TableRec { float lo, hi; uint32_t n, pad; };

// Everything a kernel needs to know about the scene: passed by value as a kernel parameter.
// The same struct points at host arrays when the stage functions are exercised by the CPU unit
// tests (tests/host_emu.cpp).
struct SceneView {
    const Node4* nodes;
    const Prim* prims;
    const TriShade* tri_shade;
    const TriFrames* tri_frames;
    const PlaneRec* planes;
    const MarchedRec* marched;
    const MaterialRec* materials;
    const ComponentRec* components;
    const ProgramRec* programs;
    const Instr* code;
    const SpectrumRec* spectra;
    const float* spectrum_data;
    const TextureRec* textures;   // colour textures first, then mono textures
    const float* texels;
    const LampRec* lamps;
    const TileRec* tiles;
    const float* burns;           // r, g, b interleaved
    const float* xyz;             // x, y, z interleaved
    const float* d65;
    uint32_t n_nodes, n_prims, n_planes, n_marched, n_lamps, n_tiles, n_color_textures;
    uint32_t vm_regs;             // registers the largest program uses (>= 1): sizes the on-chip register file
    int32_t root;                 // child code of the root: >= 0 interior node, < 0 leaf ~rank; only valid when n_prims > 0
    int32_t sky_program, filter_program, white_program;
    float root_lo[3], root_hi[3];
    TableRec burns_t, xyz_t, d65_t;
    CameraRec camera;
    RendererRec renderer;
    FilmRec film;
};

// ---- pipeline records --------------------------------------------------------------------
// 32-byte ray.  mode 0: closest hit (World::intersect, world.rs:273-299).  mode 1: visibility
// ray of tracer.rs:381-389 - only "is there a hit with DIST_EPSILON < t and t*t < limit" is
// needed, which equals the reference's closest-hit test `hit.distance^2 >= sq - eps`; mode 2: the
// same with `t < limit` (bidirectional.rs:346-351, cameras.rs:138-142).
struct alignas(32) Ray {   // one 32-byte sector: moved with single 256-bit loads / stores
    float o[3];
    uint32_t mode;
    float d[3];
    float limit;
};
struct alignas(32) Hit {  // 32 B, one sector
    float t, u, v;
    uint32_t rank;      // primitive rank, plane index, or 0xFFFFFFFF
    uint32_t kind;      // PYR_KIND_*
    uint32_t nodes, leaves, pad;  // statistics (only filled in stats mode)
};

enum : uint32_t { KIND_MISS = 0, KIND_PLANE = 1, KIND_TRIANGLE = 2, KIND_SPHERE = 3, KIND_RAY_MARCHED = 4 };

}  // namespace pyr
