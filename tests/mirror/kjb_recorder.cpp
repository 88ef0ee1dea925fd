// TEST INFRASTRUCTURE ONLY — a third "backend" of the C-ABI that computes nothing: every entry point records its name and the raw bytes of
// its argument struct.  Linked with the UNMODIFIED host mirror (kajiya_b200/csrc/host/kjb_world.cpp), it shows exactly which passes the
// mirror issues, in which order, with which resources bound in which argument slot and which constants — so tests/test_host_mirror.py can
// hold the mirror against (i) the pass table extracted from kajiya's Rust sources (tests/golden/pass_table.json) and (ii) an independent
// numpy restatement of the per-frame constants.  Resources are fake handles (never dereferenced): image n has data = 0x7f00_0000_0000 + n << 20.
#include "../../include/kjb.h"
#include <cstring>
#include <string>
#include <vector>

struct kjb_context { uint64_t launches = 0; std::string err; };
namespace {
struct Rec { std::string name; std::vector<uint8_t> bytes; };
std::vector<Rec> g_log;
uint64_t g_next = 1;
void* fake() { return (void*)(uintptr_t)(0x7f0000000000ull + (g_next++ << 20)); }
void rec(const char* name, const void* p, size_t n) { Rec r; r.name = name; if (p && n) r.bytes.assign((const uint8_t*)p, (const uint8_t*)p + n); g_log.push_back(std::move(r)); }
uint32_t texel_bytes(uint32_t f) {
    switch (f) { case 1: case 6: case 7: case 8: case 12: case 13: case 14: return 4; case 2: case 5: case 11: case 16: return 8; case 3: case 4: return 16; case 9: case 10: return 1; case 15: return 2; default: return 0; }
}
}  // namespace

extern "C" {
// ---- read-back of the log
uint32_t kjb_rec_count(void) { return uint32_t(g_log.size()); }
const char* kjb_rec_name(uint32_t i) { return g_log[i].name.c_str(); }
uint32_t kjb_rec_size(uint32_t i) { return uint32_t(g_log[i].bytes.size()); }
const void* kjb_rec_data(uint32_t i) { return g_log[i].bytes.data(); }
void kjb_rec_clear(void) { g_log.clear(); }

int kjb_abi_version(void) { return 1; }
int kjb_create(int, kjb_context** out) { *out = new kjb_context(); return 0; }
void kjb_destroy(kjb_context* c) { delete c; }
int kjb_sync(kjb_context*) { return 0; }
const char* kjb_last_error(kjb_context* c) { return c ? c->err.c_str() : ""; }
const char* kjb_backend_name(void) { return "recorder"; }
uint64_t kjb_launch_count(kjb_context* c) { return c->launches; }
void* kjb_stream(kjb_context*) { return nullptr; }
uint32_t kjb_format_texel_bytes(uint32_t f) { return texel_bytes(f); }
int kjb_image_alloc(kjb_context*, uint32_t w, uint32_t h, uint32_t layers, uint32_t fmt, kjb_image* out) { out->data = fake(); out->width = w; out->height = h; out->format = fmt; out->layers = layers; return 0; }
int kjb_image_free(kjb_context*, kjb_image*) { return 0; }
int kjb_image_clear(kjb_context*, const kjb_image* i) { rec("kjb_image_clear", i, sizeof(*i)); return 0; }
int kjb_image_copy(kjb_context*, const kjb_image* d, const kjb_image* s) { kjb_image two[2] = {*d, *s}; rec("kjb_image_copy", two, sizeof(two)); return 0; }
int kjb_image_fill_u8(kjb_context*, const kjb_image* i, uint32_t) { rec("kjb_image_fill_u8", i, sizeof(*i)); return 0; }
int kjb_image_upload(kjb_context*, const kjb_image*, const void*) { return 0; }
int kjb_image_download(kjb_context*, const kjb_image*, void*) { return 0; }
int kjb_buffer_alloc(kjb_context*, uint64_t n, kjb_buffer* out) { out->data = fake(); out->size_bytes = n; return 0; }
int kjb_buffer_free(kjb_context*, kjb_buffer*) { return 0; }
int kjb_buffer_upload(kjb_context*, const kjb_buffer*, uint64_t, const void*, uint64_t) { return 0; }
int kjb_buffer_download(kjb_context*, const kjb_buffer*, uint64_t, void*, uint64_t) { return 0; }
int kjb_timer_record(kjb_context*, uint32_t) { return 0; }
int kjb_timer_elapsed_ms(kjb_context*, uint32_t, uint32_t, float* o) { *o = 0; return 0; }
int kjb_scene_set_geometry(kjb_context*, const void*, uint64_t, const kjb_gpu_mesh*, const uint32_t*, uint32_t) { return 0; }
int kjb_scene_set_textures(kjb_context*, const kjb_texture_desc*, uint32_t) { return 0; }
int kjb_rebuild_tlas(kjb_context*, const kjb_instance* inst, uint32_t n) { rec("kjb_rebuild_tlas", inst, n * sizeof(kjb_instance)); return 0; }
int kjb_graph_begin(kjb_context*) { return 0; }
int kjb_graph_end(kjb_context*) { return 0; }
int kjb_graph_select(kjb_context*, uint32_t) { return 0; }
int kjb_set_pass_queue(kjb_context*, uint32_t q) { return q == 0 ? 0 : 1; }
int kjb_async_passes_supported(kjb_context*) { return 0; }
int kjb_graph_stats(kjb_context*, uint64_t out[2]) { out[0] = out[1] = 0; return 0; }
int kjb_tlas_stats(kjb_context*, uint64_t out[2]) { out[0] = out[1] = 0; return 0; }
int kjb_set_frame_constants(kjb_context*, const kjb_frame_constants* fc, const kjb_triangle_light*, uint32_t) { rec("kjb_set_frame_constants", fc, sizeof(*fc)); return 0; }
int kjb_set_luts(kjb_context*, const kjb_image*, const kjb_image*) { return 0; }
int kjb_ray_counters(kjb_context*, uint64_t out[2], int) { out[0] = out[1] = 0; return 0; }
int kjb_comm_nccl_unique_id(void*) { return 1; }
int kjb_comm_init_nccl(kjb_context*, const void*, uint32_t, uint32_t) { return 1; }
int kjb_comm_set_callback(kjb_context*, kjb_allgather_fn, void*, uint32_t, uint32_t) { return 0; }
int kjb_comm_rank(kjb_context*, uint32_t* r, uint32_t* n) { *r = 0; *n = 1; return 0; }
int kjb_allgather(kjb_context*, const void*, void*, uint64_t n) { rec("kjb_allgather", &n, sizeof(n)); return 0; }
int kjb_allgather_on(kjb_context*, uint32_t, const void*, void*, uint64_t n) { rec("kjb_allgather", &n, sizeof(n)); return 0; }
int kjb_memcpy_d2d(kjb_context*, void*, const void*, uint64_t) { return 0; }
int kjb_memcpy_d2d_batch(kjb_context*, const kjb_copy_desc*, uint32_t) { return 0; }
int kjb_memcpy_d2d_batch_on(kjb_context*, uint32_t, const kjb_copy_desc*, uint32_t) { return 0; }
int kjb_image_upload_on(kjb_context*, uint32_t, const kjb_image*, const void*) { return 0; }
int kjb_image_download_on(kjb_context*, uint32_t, const kjb_image*, void*) { return 0; }
int kjb_image_upload_rows_on(kjb_context*, uint32_t, const kjb_image*, const void*, uint32_t, uint32_t) { return 0; }
int kjb_image_download_rows_on(kjb_context*, uint32_t, const kjb_image*, void*, uint32_t, uint32_t) { return 0; }
int kjb_event_record(kjb_context*, uint32_t, uint32_t) { return 0; }
int kjb_queue_wait_event(kjb_context*, uint32_t, uint32_t) { return 0; }
int kjb_event_synchronize(kjb_context*, uint32_t) { return 0; }
int kjb_set_option(kjb_context*, uint32_t, uint32_t) { return 0; }
int kjb_set_scissor(kjb_context*, uint32_t y0, uint32_t y1) { uint32_t v[2] = {y0, y1}; rec("kjb_set_scissor", v, sizeof(v)); return 0; }
int kjb_set_debug_serial(kjb_context*, uint32_t) { return 0; }
// ---- every kjb_pass_* of include/kjb.h (generated from the header by tests/test_host_mirror.py)
#define REC_PASS(fn, T) int fn(kjb_context* c, const T* a) { c->launches++; rec(#fn, a, sizeof(*a)); return 0; }
#include "_build/passes.inc"
}
