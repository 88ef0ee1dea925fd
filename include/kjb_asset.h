/* kjb_asset.h — the asset side of the path ("next" row N5): glTF 2.0 scene -> TriangleMesh, encoded image -> RGBA8 mip chain.
 *
 * Mirrors crates/lib/kajiya-asset: `LoadGltfScene` (src/mesh.rs:264-441), `load_gltf_material` (src/mesh.rs:120-262),
 * the buffer/image resolution of src/import_gltf.rs:32-160, `LoadImage` (src/image.rs:62-98) and the uncompressed branch of
 * `CreateGpuImage::process_rgba8` (src/image.rs:130-283: 2048 clamp, Lanczos3 mip chain, channel swizzle).
 * It is host-only code (the reference's is too: an offline bake step) in its own library, libkjb_asset.so; the TriangleMesh it
 * returns is laid out as the `kjb_mesh_desc` that kjb_world_add_mesh (WorldRenderer::add_mesh) consumes, so the two calls chain
 * without a copy.  DDS images (`process_dds`, image.rs:285-335: DX10-header BC1_SRGB / BC3 / BC5 files with their own mips) are decoded
 * block by block into the same RGBA8 chains.  Not mirrored: BC5/BC7 block COMPRESSION (intel_tex_2 ISPC encoders; textures stay RGBA8, which is
 * what the hit shader's SampleLevel sees after hardware decode up to the encoder's loss), mikktspace tangent generation
 * (tangents only feed the raster normal-map path; the ray-traced hit shader has normal mapping compiled out, gbuffer.rchit.hlsl:124-167).
 */
#ifndef KJB_ASSET_H
#define KJB_ASSET_H
#include "kjb_world.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kjb_asset kjb_asset;

/* LoadGltfScene { path, scale, rotation } (mesh.rs:264-269).  rotation = quaternion xyzw (NULL = identity).
 * .gltf (JSON + external / data: buffers) and .glb containers; default scene or the first one; every node's mesh primitives with
 * POSITION and NORMAL are appended in traversal order, pre-transformed by the node's world matrix; one material (and its four maps:
 * normal, spec, albedo, emissive) per primitive. Returns 0 on success; on failure *out = NULL and kjb_asset_last_error() says why. */
int  kjb_asset_load_gltf(const char *path, float scale, const float rotation_xyzw[4], kjb_asset **out);
void kjb_asset_destroy(kjb_asset *a);
/* thread-local message of the last failed call */
const char *kjb_asset_last_error(void);

/* TriangleMesh as `kjb_mesh_desc` views into the asset (valid until kjb_asset_destroy).  Image maps are decoded and mip-chained at
 * load time with their TexParams (albedo/emissive sRGB, spec swizzled [1,2,0,3], mips on).  use_lights is left 0. */
int  kjb_asset_get_mesh(const kjb_asset *a, kjb_mesh_desc *out);
/* TriangleMesh::tangents (4 per vertex): the file's TANGENT attribute transformed like the reference does, else (1,0,0,0) */
const float *kjb_asset_tangents(const kjb_asset *a);
/* counts for reports: [0] nodes visited, [1] primitives appended, [2] primitives skipped (no POSITION/NORMAL), [3] images decoded */
int  kjb_asset_stats(const kjb_asset *a, uint32_t out[4]);

/* LoadImage (image.rs:62-98): PNG (all colour types / bit depths, Adam7), baseline + progressive JPEG, DX10 DDS (top level) -> tightly packed RGBA8.
 * The caller frees *out_rgba8 with kjb_asset_free_buffer. */
int  kjb_asset_decode_image(const uint8_t *bytes, uint64_t byte_count, uint8_t **out_rgba8, uint32_t *out_width, uint32_t *out_height);
/* CreateGpuImage::process_rgba8 with TexCompressionMode::None (image.rs:130-283): clamp to 2048 with Lanczos3, full mip chain by
 * repeated Lanczos3 halving (each level from the previous UNswizzled level), optional channel swizzle (NULL = none) applied per level.
 * Output: mips stored one after another, level 0 first; *out_mip_count levels; (*out_width, *out_height) = level-0 extent. */
int  kjb_asset_build_mips(const uint8_t *rgba8, uint32_t width, uint32_t height, uint32_t use_mips, const uint32_t channel_swizzle[4],
                          uint8_t **out_texels, uint64_t *out_bytes, uint32_t *out_width, uint32_t *out_height, uint32_t *out_mip_count);
void kjb_asset_free_buffer(void *p);

#ifdef __cplusplus
}
#endif
#endif
