// jb_api.cu -- the extern "C" surface declared in include/jpegb200.h.
#include <string.h>
#include <atomic>

#include "jb_common.cuh"
#include "jb_forward.cuh"
#include "jb_inverse.cuh"

#define JB_CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return jb_cuda_fail(_e); } while (0)

static int jb_cuda_fail(cudaError_t e) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice)
        return JB_ERR_NO_DEVICE;
    return JB_ERR_CUDA;
}

extern "C" int jb_version(void) { return JB_VERSION; }

extern "C" const char* jb_strerror(int code) {
    switch (code) {
    case JB_OK: return "ok";
    case JB_ERR_BAD_PARAM: return "bad parameter";
    case JB_ERR_BAD_QUANTIZATION: return "bad quantization (BadQuantizationError)";
    case JB_ERR_EMPTY_ARRAY: return "empty array (EmptyArrayError)";
    case JB_ERR_UNSUPPORTED: return "dct_size or block_size beyond the supported range";
    case JB_ERR_WORKSPACE: return "workspace too small";
    case JB_ERR_OUT_CAPACITY: return "output buffer too small";
    case JB_ERR_CUDA: return "CUDA runtime error";
    case JB_ERR_BAD_RLE_CODE: return "amplitude does not fit a 15-bit size field (BadRleCodeError)";
    case JB_ERR_BAD_STREAM: return "malformed stream";
    case JB_ERR_NO_DEVICE: return "no usable CUDA device";
    default: return "unknown error";
    }
}

extern "C" int jb_geometry_of(const jb_params* p, jb_geometry* geo) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!geo) return JB_ERR_BAD_PARAM;
    geo->h1 = g.H1; geo->w1 = g.W1; geo->h2 = g.H2; geo->w2 = g.W2;
    geo->vb = g.vb; geo->hb = g.hb;
    geo->blocks_per_plane = g.nblocks;
    geo->max_block_bytes = g.maxblk;
    geo->chunks_per_plane = g.cpp;
    geo->reserved = 0;
    return JB_OK;
}

extern "C" size_t jb_max_stream_bytes(const jb_params* p, int n_planes) {
    JbGeom g;
    if (jb_make_geom(p, &g) != JB_OK || n_planes <= 0) return 0;
    return (size_t)n_planes * (size_t)g.nblocks * (size_t)g.maxblk;
}

// ---- compress ----------------------------------------------------------------------------------
// the chunk size of this call (jb_call_chunk_blocks): geometry as the kernels of the call see it
static JbGeom jb_geom_for_call(const JbGeom& g0, int n_planes) {
    JbGeom g = g0;
    g.chunk = jb_call_chunk_blocks(g.d, (long long)n_planes * g.nblocks);
    g.cpp = (g.nblocks + g.chunk - 1) / g.chunk;
    return g;
}

struct JbFwdWs { size_t ctrl, chunk_len, chunk_off, seg_total, tmp_small, tmp, total; unsigned chunk_cap; };
static JbFwdWs jb_fwd_ws(int d, int chunk, size_t n_chunks) {
    JbFwdWs w;
    size_t o = jb_align_up(jb_table_layout(d).total, 256);
    w.chunk_cap = (unsigned)jb_align_up((size_t)chunk * jb_max_block_bytes(d * d) + 32, 16);
    if (w.chunk_cap < 1024 + 32) w.chunk_cap = 1024 + 32;      // the gather kernel reads 1 KB ahead
    w.ctrl = o;      o += JB_CTRL_BYTES;
    w.chunk_len = o; o += jb_align_up(n_chunks * 4, 256);
    w.chunk_off = o; o += jb_align_up(n_chunks * 4, 256);
    w.seg_total = o; o += jb_align_up(((n_chunks + JB_SCAN_SEG - 1) / JB_SCAN_SEG) * 8, 256);
    w.tmp_small = o; o += jb_align_up(n_chunks * (size_t)JB_SLOT_STRIDE, 256);
    w.tmp = o;       o += jb_align_up(n_chunks * (size_t)w.chunk_cap, 256);
    w.total = o;
    return w;
}

extern "C" size_t jb_compress_workspace_bytes(const jb_params* p, int n_planes) {
    JbGeom g;
    if (jb_make_geom(p, &g) != JB_OK || n_planes <= 0) return 0;
    g = jb_geom_for_call(g, n_planes);
    return jb_fwd_ws(g.d, g.chunk, (size_t)n_planes * g.cpp).total;
}

extern "C" size_t jb_stage_pack_workspace_bytes(int n_planes, int blocks_per_plane, int dct_size) {
    if (n_planes <= 0 || blocks_per_plane <= 0 || dct_size < 1 || dct_size > JB_MAX_DCT_SIZE) return 0;
    const size_t cb = (size_t)jb_chunk_blocks(dct_size);
    size_t cpp = ((size_t)blocks_per_plane + cb - 1) / cb;
    return jb_fwd_ws(dct_size, (int)cb, (size_t)n_planes * cpp).total;
}

// The status block of calls that do not run the self-cleaning protocol (word 1 = "no bad code yet" = all ones).
__global__ void jb_init_kernel(unsigned long long* status, unsigned* ticket) {
    const int t = threadIdx.x;
    if (t < JB_STATUS_WORDS) status[t] = t == 1 ? ~0ull : 0ull;
    if (ticket && t < 64) ticket[t] = 0u;
}

static int jb_reset_status(uint64_t* d_status, unsigned* d_ticket, cudaStream_t s) {
    JB_LAUNCH((jb_init_kernel), 1, 64, 0, s, (unsigned long long*)d_status, d_ticket);
    JB_CUDA_TRY(cudaGetLastError());
    return JB_OK;
}

static int jb_forward_common(int mode, const uint8_t* d_planes, size_t plane_stride, size_t row_pitch,
                             int n_planes, const JbGeom& g_in, uint8_t* d_out, size_t out_cap,
                             uint64_t* d_plane_off, uint64_t* d_status, int16_t* d_coeffs_out,
                             const int32_t* d_coeffs_in, void* d_ws, size_t ws_bytes, cudaStream_t s) {
    const JbGeom g = mode == 2 ? g_in : jb_geom_for_call(g_in, n_planes);       // (mode 2: the caller built g by hand)
    const size_t n_chunks = (size_t)n_planes * g.cpp;
    if (n_chunks > 0x7FFFFFFFull) return JB_ERR_UNSUPPORTED;
    JbFwdWs w = jb_fwd_ws(g.d, g.chunk, n_chunks);
    if (!d_ws || ws_bytes < w.total) return JB_ERR_WORKSPACE;
    if (((uintptr_t)d_ws & 255) != 0) return JB_ERR_BAD_PARAM;
    char* ws = (char*)d_ws;
    // Control block protocol (jb_forward.cuh): a call without JB_FLAG_REUSE_TABLES establishes the clean state
    // together with the tables; every call leaves it clean again.  Two launches per call after the first one:
    // the fused transform kernel and the gather.
    if (!(g.flags & JB_FLAG_REUSE_TABLES))
        JB_CUDA_TRY(jb_launch_init_ctrl((unsigned*)(ws + w.ctrl), nullptr, 0, s));

    JbFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.t = jb_tables_at(d_ws, g.d);
    a.planes = d_planes; a.plane_stride = plane_stride; a.row_pitch = row_pitch;
    a.n_planes = n_planes; a.n_chunks = (unsigned)n_chunks;
    a.out = d_out; a.out_cap = out_cap;
    a.plane_off = (unsigned long long*)d_plane_off;
    a.ctrl = (unsigned*)(ws + w.ctrl);
    a.status_out = (unsigned long long*)d_status;
    a.chunk_len = (unsigned*)(ws + w.chunk_len);
    a.chunk_off = (unsigned*)(ws + w.chunk_off);
    a.seg_total = (unsigned long long*)(ws + w.seg_total);
    a.tmp = (uint8_t*)(ws + w.tmp);
    a.tmp_small = (uint8_t*)(ws + w.tmp_small);
    a.chunk_cap = w.chunk_cap;
    a.coeffs_out = d_coeffs_out; a.coeffs_in = d_coeffs_in;
    if (mode != 2 && !(g.flags & JB_FLAG_REUSE_TABLES)) JB_CUDA_TRY(jb_launch_build_tables(g, a.t, s));
    if (mode == 0) jb_prof_mark(0, s);                 // (the matching end mark sits in jb_launch_scan_gather)
    if (mode != 2 && !(g.flags & JB_FLAG_FORCE_GENERIC) && jb_fwd_fast_eligible(g))
        JB_CUDA_TRY(jb_launch_fwd_fast(a, mode, s));
    else if (mode != 2 && !(g.flags & JB_FLAG_FORCE_GENERIC) && jb_fwd_large_eligible(g))
        JB_CUDA_TRY(jb_launch_fwd_large(a, mode, s));
    else
        JB_CUDA_TRY(jb_launch_fwd_generic(a, mode, s));
    if (mode == 1) JB_CUDA_TRY(jb_launch_finish(a, s));      // (modes 0 and 2 end with the gather, which does this)
    return JB_OK;
}

extern "C" int jb_compress_planes(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                                  const jb_params* p, uint8_t* d_out, size_t out_cap, uint64_t* d_plane_off,
                                  uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!d_planes || !d_out || !d_plane_off || !d_status || n_planes <= 0) return JB_ERR_BAD_PARAM;
    if (row_pitch < (size_t)g.W || plane_stride < row_pitch * (size_t)(g.H - 1) + (size_t)g.W) return JB_ERR_BAD_PARAM;
    return jb_forward_common(0, d_planes, plane_stride, row_pitch, n_planes, g, d_out, out_cap, d_plane_off,
                             d_status, nullptr, nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int jb_stage_forward_coeffs(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                                       const jb_params* p, int16_t* d_coeffs, uint64_t* d_status,
                                       void* d_ws, size_t ws_bytes, void* stream) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!d_planes || !d_coeffs || !d_status || n_planes <= 0) return JB_ERR_BAD_PARAM;
    if (row_pitch < (size_t)g.W || plane_stride < row_pitch * (size_t)(g.H - 1) + (size_t)g.W) return JB_ERR_BAD_PARAM;
    return jb_forward_common(1, d_planes, plane_stride, row_pitch, n_planes, g, nullptr, 0, nullptr, d_status,
                             d_coeffs, nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int jb_stage_pack(const int32_t* d_coeffs, int n_planes, int blocks_per_plane, int dct_size,
                             uint8_t* d_out, size_t out_cap, uint64_t* d_plane_off, uint64_t* d_status,
                             void* d_ws, size_t ws_bytes, void* stream) {
    if (!d_coeffs || !d_out || !d_plane_off || !d_status || n_planes <= 0 || blocks_per_plane <= 0)
        return JB_ERR_BAD_PARAM;
    if (dct_size < 1) return JB_ERR_BAD_PARAM;
    if (dct_size > JB_MAX_DCT_SIZE) return JB_ERR_UNSUPPORTED;
    JbGeom g;
    memset(&g, 0, sizeof(g));
    g.d = dct_size; g.n = dct_size * dct_size; g.nblocks = blocks_per_plane;
    g.maxblk = jb_max_block_bytes(g.n);
    g.chunk = jb_chunk_blocks(dct_size);
    g.cpp = (blocks_per_plane + g.chunk - 1) / g.chunk;
    g.bs = 1;
    return jb_forward_common(2, nullptr, 0, 0, n_planes, g, d_out, out_cap, d_plane_off, d_status, nullptr,
                             d_coeffs, d_ws, ws_bytes, (cudaStream_t)stream);
}

cudaError_t jb_launch_stage_f64(const JbGeom& g, const JbTables& t, const uint8_t* d_planes, size_t plane_stride,
                                size_t row_pitch, int n_planes, int which, double* d_out, cudaStream_t s);

extern "C" int jb_stage_float64(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                                const jb_params* p, int which, double* d_out, void* d_ws, size_t ws_bytes, void* stream) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!d_planes || !d_out || n_planes <= 0 || n_planes > 65535) return JB_ERR_BAD_PARAM;
    if (which != JB_F64_SAMPLES && which != JB_F64_TRANSFORM && which != JB_F64_PREROUNDING) return JB_ERR_BAD_PARAM;
    if (row_pitch < (size_t)g.W || plane_stride < row_pitch * (size_t)(g.H - 1) + (size_t)g.W) return JB_ERR_BAD_PARAM;
    if (!d_ws || ws_bytes < jb_table_layout(g.d).total) return JB_ERR_WORKSPACE;
    if (((uintptr_t)d_ws & 255) != 0) return JB_ERR_BAD_PARAM;
    cudaStream_t s = (cudaStream_t)stream;
    const JbTables t = jb_tables_at(d_ws, g.d);
    if (!(g.flags & JB_FLAG_REUSE_TABLES)) JB_CUDA_TRY(jb_launch_build_tables(g, t, s));
    JB_CUDA_TRY(jb_launch_stage_f64(g, t, d_planes, plane_stride, row_pitch, n_planes, which, d_out, s));
    return JB_OK;
}

// ---- measurement hooks -----------------------------------------------------------------------------
static std::atomic<unsigned long long> g_jb_launches{0};
void jb_note_launches(unsigned n) { g_jb_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" unsigned long long jb_debug_launch_count(void) { return g_jb_launches.load(std::memory_order_relaxed); }

static cudaEvent_t g_prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};

extern "C" int jb_debug_kernel_events(void* fwd_begin, void* fwd_end, void* inv_begin, void* inv_end) {
    if ((fwd_begin == nullptr) != (fwd_end == nullptr) || (inv_begin == nullptr) != (inv_end == nullptr)) return JB_ERR_BAD_PARAM;
    g_prof_ev[0] = (cudaEvent_t)fwd_begin; g_prof_ev[1] = (cudaEvent_t)fwd_end;
    g_prof_ev[2] = (cudaEvent_t)inv_begin; g_prof_ev[3] = (cudaEvent_t)inv_end;
    return JB_OK;
}

void jb_prof_mark(int which, cudaStream_t s) {
    if (g_prof_ev[which]) cudaEventRecord(g_prof_ev[which], s);
}

// ---- decompress --------------------------------------------------------------------------------
extern "C" size_t jb_decompress_workspace_bytes(const jb_params* p, int n_planes, size_t in_bytes) {
    JbGeom g;
    if (jb_make_geom(p, &g) != JB_OK || n_planes <= 0) return 0;
    return jb_dec_layout(g.d, n_planes, g.nblocks, in_bytes, jb_table_layout(g.d).total).total;
}

extern "C" size_t jb_stage_unpack_workspace_bytes(int n_planes, int blocks_per_plane, int dct_size, size_t in_bytes) {
    if (n_planes <= 0 || blocks_per_plane <= 0 || dct_size < 1 || dct_size > JB_MAX_DCT_SIZE) return 0;
    return jb_dec_layout(dct_size, n_planes, blocks_per_plane, in_bytes, jb_table_layout(dct_size).total).total;
}

extern "C" int jb_decompress_framing_path(const jb_params* p, int n_planes, size_t in_bytes) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (n_planes <= 0) return JB_ERR_BAD_PARAM;
    JbDecLayout L = jb_dec_layout(g.d, n_planes, g.nblocks, in_bytes, jb_table_layout(g.d).total);
    return jb_framing_is_chain(L.max_tiles, n_planes) ? JB_FRAMING_CHAIN : JB_FRAMING_STITCH;
}

static int jb_inverse_common(int mode, const uint8_t* d_in, size_t in_bytes, const uint64_t* d_plane_off,
                             const uint64_t* d_plane_len, int n_planes, const JbGeom& g_in,
                             uint8_t* d_planes_out, size_t plane_stride, size_t row_pitch,
                             int16_t* d_coeffs_out, const int16_t* d_coeffs_in, uint64_t* d_status,
                             void* d_ws, size_t ws_bytes, cudaStream_t s) {
    const JbGeom g = mode == 1 ? g_in : jb_geom_for_call(g_in, n_planes);       // (mode 1: the caller built g by hand)
    const size_t n_chunks = (size_t)n_planes * g.cpp;
    if (n_chunks > 0x7FFFFFFFull) return JB_ERR_UNSUPPORTED;
    const size_t table_bytes = jb_table_layout(g.d).total;
    if (!d_ws || ((uintptr_t)d_ws & 255) != 0) return JB_ERR_WORKSPACE;
    char* ws = (char*)d_ws;
    if (mode == 2) {                                  // (with framing, jb_frame_prep_kernel resets the status block)
        int rc = jb_reset_status(d_status, nullptr, s);
        if (rc != JB_OK) return rc;
    }

    JbInvArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.t = jb_tables_at(d_ws, g.d);
    a.n_planes = n_planes; a.n_chunks = (unsigned)n_chunks;
    a.planes_out = d_planes_out; a.plane_stride = plane_stride; a.row_pitch = row_pitch;
    a.coeffs_out = d_coeffs_out; a.coeffs_in = d_coeffs_in;
    a.status = (unsigned long long*)d_status;          // (mode 2: no framing, no control block: errors go straight out)

    if (mode != 2) {
        JbDecLayout L = jb_dec_layout(g.d, n_planes, g.nblocks, in_bytes, table_bytes);
        if (ws_bytes < L.total) return JB_ERR_WORKSPACE;
        // control block protocol (jb_common.cuh): clean when the call starts, left clean by its last kernel
        unsigned* ctrl = (unsigned*)(ws + L.ctrl);
        if (!(g.flags & JB_FLAG_REUSE_TABLES)) JB_CUDA_TRY(jb_launch_init_ctrl(ctrl, nullptr, 0, s));
        a.status = (unsigned long long*)(ws + L.ctrl + JB_CTRL_STATUS_OFF);
        a.status_out = (unsigned long long*)d_status;
        a.ticket = ctrl;
        a.done = ctrl + 1;
        JbFrameArgs f;
        memset(&f, 0, sizeof(f));
        f.in = d_in;
        f.in_bytes = in_bytes;
        f.plane_off = (const unsigned long long*)d_plane_off;
        f.plane_len = (const unsigned long long*)d_plane_len;
        f.n_planes = n_planes; f.n = g.n; f.nblocks = g.nblocks; f.maxblk = g.maxblk;
        f.max_tiles = L.max_tiles; f.tile_bytes = L.tile_bytes;
        f.force_serial = (g.flags & JB_FLAG_SERIAL_FRAMING) ? 1 : 0;
        f.pdl = (g.flags & JB_FLAG_PDL) ? 1 : 0;
        f.tile_first = (unsigned*)(ws + L.tile_first);
        f.fallback = (unsigned*)(ws + L.fallback);
        f.notplain = (unsigned*)(ws + L.notplain);
        f.big_list = (unsigned*)(ws + L.big_list);
        f.block_start = (unsigned*)(ws + L.block_start);
        f.tile_exit = (unsigned*)(ws + L.tile_exit);
        f.tile_entry = (unsigned*)(ws + L.tile_entry);
        f.tile_from = (unsigned*)(ws + L.tile_from);
        f.tile_npriv = (unsigned*)(ws + L.tile_npriv);
        f.tile_hops = (unsigned*)(ws + L.tile_hops);
        f.tile_base = (unsigned*)(ws + L.tile_base);
        f.vbits = (uint32_t*)(ws + L.vbits);
        f.warp_first = (unsigned*)(ws + L.warp_first);
        f.warp_stream = (unsigned*)(ws + L.warp_stream);
        f.status = a.status;
        // (a stream that fails framing leaves its part of block_start unwritten; the call then reports
        // JB_ERR_BAD_STREAM and the transform kernels check every offset they read against the stream bounds)
        JB_CUDA_TRY(jb_launch_framing(f, s));
        a.in = d_in; a.in_bytes = in_bytes;
        a.plane_off = f.plane_off; a.plane_len = f.plane_len;
        a.block_start = f.block_start;
    } else if (ws_bytes < table_bytes) {
        return JB_ERR_WORKSPACE;
    }
    if (mode != 1 && !(g.flags & JB_FLAG_REUSE_TABLES)) JB_CUDA_TRY(jb_launch_build_tables(g, a.t, s));
    if (mode == 0) jb_prof_mark(2, s);
    if (mode != 1 && !(g.flags & JB_FLAG_FORCE_GENERIC) && jb_inv_fast_eligible(g)) {
        JB_CUDA_TRY(jb_launch_inv_fast(a, mode, s));      // (publishes the status and cleans the control block itself)
        if (mode == 0) jb_prof_mark(3, s);
    } else if (mode != 1 && !(g.flags & JB_FLAG_FORCE_GENERIC) && jb_inv_large_eligible(g)) {
        JB_CUDA_TRY(jb_launch_inv_large(a, mode, s));     // (likewise)
        if (mode == 0) jb_prof_mark(3, s);
    } else {
        {
            JB_CUDA_TRY(jb_launch_inv_generic(a, mode, s));
        }
        if (mode == 0) jb_prof_mark(3, s);
        if (a.status_out) JB_CUDA_TRY(jb_launch_dec_finish(a, s));
    }
    return JB_OK;
}

extern "C" int jb_decompress_planes(const uint8_t* d_in, size_t in_bytes, const uint64_t* d_plane_off,
                                    const uint64_t* d_plane_len, int n_planes, const jb_params* p,
                                    uint8_t* d_planes_out, size_t plane_stride, size_t row_pitch,
                                    uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!d_in || !d_plane_off || !d_plane_len || !d_planes_out || !d_status || n_planes <= 0) return JB_ERR_BAD_PARAM;
    if (row_pitch < (size_t)g.W || plane_stride < row_pitch * (size_t)(g.H - 1) + (size_t)g.W) return JB_ERR_BAD_PARAM;
    return jb_inverse_common(0, d_in, in_bytes, d_plane_off, d_plane_len, n_planes, g, d_planes_out, plane_stride,
                             row_pitch, nullptr, nullptr, d_status, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int jb_stage_unpack(const uint8_t* d_in, size_t in_bytes, const uint64_t* d_plane_off,
                               const uint64_t* d_plane_len, int n_planes, int blocks_per_plane, int dct_size, int flags,
                               int16_t* d_coeffs, uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream) {
    if (!d_in || !d_plane_off || !d_plane_len || !d_coeffs || !d_status || n_planes <= 0 || blocks_per_plane <= 0)
        return JB_ERR_BAD_PARAM;
    if (dct_size < 1) return JB_ERR_BAD_PARAM;
    if (dct_size > JB_MAX_DCT_SIZE) return JB_ERR_UNSUPPORTED;
    JbGeom g;
    memset(&g, 0, sizeof(g));
    g.d = dct_size; g.n = dct_size * dct_size; g.nblocks = blocks_per_plane;
    g.maxblk = jb_max_block_bytes(g.n);
    g.chunk = jb_chunk_blocks(dct_size);
    g.cpp = (blocks_per_plane + g.chunk - 1) / g.chunk;
    g.bs = 1;
    g.flags = flags & JB_FLAG_SERIAL_FRAMING;
    return jb_inverse_common(1, d_in, in_bytes, d_plane_off, d_plane_len, n_planes, g, nullptr, 0, 0, d_coeffs,
                             nullptr, d_status, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int jb_stage_inverse_coeffs(const int16_t* d_coeffs, int n_planes, const jb_params* p,
                                       uint8_t* d_planes_out, size_t plane_stride, size_t row_pitch,
                                       uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream) {
    JbGeom g;
    int rc = jb_make_geom(p, &g);
    if (rc != JB_OK) return rc;
    if (!d_coeffs || !d_planes_out || !d_status || n_planes <= 0) return JB_ERR_BAD_PARAM;
    if (row_pitch < (size_t)g.W || plane_stride < row_pitch * (size_t)(g.H - 1) + (size_t)g.W) return JB_ERR_BAD_PARAM;
    return jb_inverse_common(2, nullptr, 0, nullptr, nullptr, n_planes, g, d_planes_out, plane_stride, row_pitch,
                             nullptr, d_coeffs, d_status, d_ws, ws_bytes, (cudaStream_t)stream);
}
