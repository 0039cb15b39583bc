/* pyrite_b200 - layout of the project IR blob that pyr_project_load() takes.
 *
 * The blob is what pyrite holds after `load_project` (pyrite/src/project/mod.rs:29-93) and before `parse_project`
 * (main.rs:111-134): the typed `Project` (project/mod.rs:103-203) with its expression nodes (project/expressions.rs:152-201),
 * surface-material nodes (project/materials.rs:7-34), spectra (project/spectra.rs:14-28), decoded textures
 * (project/textures.rs:68-113) and parsed meshes (project/meshes.rs:54-131), plus the three constant tables the reference
 * compiles in from pyrite/data/*.csv (build.rs).  A host in any language writes it with the rules below; the two writers
 * that exist are pyrite_b200/project.py (`serialize_project`) and the C++ reader pyrite_b200/csrc/project_ir.hpp is the
 * normative decoder (it rejects anything malformed with PYR_ERR_INVALID).  INTEGRATION.md shows the Rust side.
 *
 * Conventions: little-endian; every field is a multiple of 4 bytes; u32 / i32 / f32 / f64 are unpadded.
 *
 *   ex        := u32 tag ; tag == 0: f64 number              (Expression::Number,  expressions.rs:65-71)
 *                          tag != 0: u32 node, u32 0          (Expression::Complex: index into the expression-node table)
 *   opt<T>    := u32 present (0 | 1) ; T if present           (Option<T>)
 *   opt_u32   := u32 present ; u32 value (always written)
 *   text      := u32 n ; n bytes ; zero padding to a multiple of 4
 *   look_at   := ex from ; ex to ; opt<ex> up                 (Transform::LookAt, project/mod.rs:255-269)
 *   material  := u32 surface (index into the surface table) ; opt<ex> normal_map      (project/mod.rs:216-220)
 *
 *   blob :=
 *     u32 PYR_IR_MAGIC ; u32 PYR_IR_VERSION
 *     u32 n_nodes ; n_nodes x node                            nodes may only be referenced by index; identity = one node
 *         node := u32 kind (pyr_ir_node) ; [u32 operator (pyr_ir_binop) if kind == BINARY] ;
 *                 [u32 resource if kind >= SPECTRUM: index into spectra / colour textures / mono textures] ;
 *                 arity(kind) x ex      VECTOR: x y z w | RGB: red green blue | BINARY: lhs rhs | MIX: amount lhs rhs |
 *                                       CLAMP: value min max | FRESNEL: ior env_ior | BLACKBODY: temperature | others: none
 *     u32 n_surfaces ; n_surfaces x surface
 *         surface := u32 kind (pyr_ir_surface) ;
 *                    EMISSIVE | DIFFUSE | MIRROR: ex color
 *                    REFRACTIVE: ex color ; ex ior ; opt<ex> dispersion ; opt<ex> env_ior ; opt<ex> env_dispersion
 *                    MIX: u32 lhs ; u32 rhs ; ex amount        (`amount` weights lhs, materials/mod.rs:176-194)
 *                    ADD: u32 lhs ; u32 rhs
 *     u32 n_spectra ; n_spectra x spectrum
 *         spectrum := u32 0 ; f32 min ; f32 max ; u32 n ; n x f32            (Spectrum::Array;  light_source.d65 / .a are arrays)
 *                   | u32 1 ; u32 n ; n x (f32 wavelength, f32 value)         (Spectrum::Curve)
 *     u32 n_color_textures ; each: u32 width ; u32 height ; width*height x (f32 r, g, b, a)   linear RGBA, row-major from the top
 *     u32 n_mono_textures  ; each: u32 width ; u32 height ; width*height x f32                 linear luma
 *     u32 n_meshes ; each:
 *         u32 n_positions ; 3 f32 each | u32 n_uvs ; 2 f32 each | u32 n_normals ; 3 f32 each
 *         u32 n_objects ; each: text name ; u32 n_triangles ; n_triangles x 3 corners x (i32 position, i32 uv or -1, i32 normal or -1)
 *             objects in file order, then groups, then polygons; only 3-corner polygons (world.rs:190-234)
 *     f32 min ; f32 max ; u32 n ; n x (f32 r, g, b)            Burns RGB basis        (data/srgb_cie1931.csv, build.rs:18-59: max = min + n)
 *     f32 min ; f32 max ; u32 n ; n x (f32 x, y, z)            CIE XYZ responses      (data/ciexyz65_1.csv)
 *     f32 min ; f32 max ; u32 n ; n x f32                      D65                    (data/d65.csv)
 *     u32 width ; u32 height ; opt<ex> filter ; opt<ex> white                          image       (project/mod.rs:111-118)
 *     u32 renderer (pyr_ir_renderer) ; u32 pixel_samples ;
 *       opt_u32 threads, bounces, light_samples, spectrum_samples, spectrum_resolution, tile_size, light_bounces   (mod.rs:131-161)
 *     look_at camera_transform ; ex fov ; opt<ex> focus_distance ; opt<ex> aperture    camera.perspective (mod.rs:120-129)
 *     opt<ex> sky
 *     u32 n_objects ; n_objects x object                                                world.objects, in project order
 *         object := u32 kind (pyr_ir_object) ;
 *             SPHERE: ex position ; ex radius ; opt<ex> texture_scale ; material
 *             PLANE:  ex origin ; ex normal ; opt<ex> texture_scale ; material
 *             RAY_MARCHED: u32 estimator (0 mandelbulb | 1 quaternion julia) ;
 *                 mandelbulb: ex iterations ; ex threshold ; ex power ; opt<ex> constant
 *                 julia:      ex iterations ; ex threshold ; ex constant ; ex slice_plane ; u32 variant (0 regular | 1 cubic | 2 bicomplex)
 *                 u32 bounds (0 box | 1 sphere) ; ex a (box min | sphere position) ; ex b (box max | sphere radius) ; material
 *             MESH: u32 mesh ; u32 n ; n x (text object_name ; material) ; opt<ex> scale ; u32 has_transform ; [look_at transform]
 *             DIRECTIONAL_LIGHT: ex direction ; ex width ; ex color
 *             POINT_LIGHT: ex position ; ex color
 *   Nothing may follow the last object.
 */
#ifndef PYRITE_B200_IR_H
#define PYRITE_B200_IR_H

#include <stdint.h>

#define PYR_IR_MAGIC 0x52495950u /* "PYIR" */
#define PYR_IR_VERSION 1u

enum pyr_ir_node { PYR_IR_VECTOR = 0, PYR_IR_RGB, PYR_IR_BINARY, PYR_IR_MIX, PYR_IR_CLAMP, PYR_IR_FRESNEL, PYR_IR_BLACKBODY,
                   PYR_IR_SPECTRUM, PYR_IR_COLOR_TEXTURE, PYR_IR_MONO_TEXTURE };
enum pyr_ir_binop { PYR_IR_ADD = 0, PYR_IR_SUB, PYR_IR_MUL, PYR_IR_DIV };
enum pyr_ir_surface { PYR_IR_EMISSIVE = 0, PYR_IR_DIFFUSE, PYR_IR_MIRROR, PYR_IR_REFRACTIVE, PYR_IR_SURFACE_MIX, PYR_IR_SURFACE_ADD };
enum pyr_ir_renderer { PYR_IR_SIMPLE = 0, PYR_IR_BIDIRECTIONAL, PYR_IR_PHOTON_MAPPING /* refused by pyr_project_load */ };
enum pyr_ir_object { PYR_IR_SPHERE = 0, PYR_IR_PLANE, PYR_IR_RAY_MARCHED, PYR_IR_MESH, PYR_IR_DIRECTIONAL_LIGHT, PYR_IR_POINT_LIGHT };

#endif /* PYRITE_B200_IR_H */
