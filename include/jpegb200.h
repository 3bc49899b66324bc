/*
 * jpegb200.h -- C ABI of libjpegb200.so: the per-block codec hot path of
 * X-rayLaser/Implementing-JPEG-compression as sm_100a CUDA kernels.
 *
 * The reference has no FFI; its boundary for this path is the pair of Python
 * functions
 *     compress_band(a, config)   -> bytes      pipeline/__init__.py:71-76
 *     decompress_band(b, config) -> ndarray    pipeline/__init__.py:79-88
 * called three times per image by Jpeg.compress / Jpeg.decompress
 * (pipeline/__init__.py:102-124).  The entry points below are what a ctypes
 * binding of that boundary calls; each one names the
 * reference stages it replaces.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer on the current CUDA device; the
 *    caller owns every buffer, including the workspace; the library allocates
 *    nothing on the device and keeps no per-call state.  Process-wide state is limited to
 *    two read-mostly caches behind a lock (the host-evaluated DCT matrix of the last
 *    dct_size, the driver entry point of the tensor-map encoder) and the measurement hook
 *    jb_debug_kernel_events, which is off unless a profiler arms it;
 *  - calls are asynchronous and ordered on `stream` (a cudaStream_t passed as
 *    void*; NULL = the legacy default stream).  Host-detectable problems return
 *    a negative code at once; problems only the device can see (an amplitude
 *    that does not fit a 15-bit size field, a malformed stream) are written to
 *    the status words the call fills, to be read after the stream is synced;
 *  - a "plane" is one colour band: height x width uint8 samples, `row_pitch`
 *    bytes between rows, `plane_stride` bytes between consecutive planes.  All
 *    planes of one call share one jb_params (an image is three planes).
 *  - streams are bit-identical to the reference's RleBytestream output
 *    (pipeline/rle_byte_stream.py:48-58) for identical quantised coefficients.
 */
#ifndef JPEGB200_H
#define JPEGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_VERSION 100

/* --transform (compress.py:39-40, pipeline/basis_change.py:15-25) */
#define JB_TRANSFORM_DCT 0
#define JB_TRANSFORM_DFT 1   /* real part of fft2 only, see SURVEY.md section 0.2 */

/* --quantization (compress.py:42-62, quantizers.py:4-53) */
#define JB_Q_NONE    0       /* RoundingQuantizer      quantizers.py:4-9   */
#define JB_Q_DISCARD 1       /* DiscardingQuantizer    quantizers.py:12-20, qparam = keep    */
#define JB_Q_DIVIDE  2       /* DivisionQuantizer      quantizers.py:23-31, qparam = divisor */
#define JB_Q_QTABLE  3       /* JpegQuantizationTable  quantizers.py:34-53, dct_size must be 8 */

/* return / status codes; the Python shim maps them to the reference's exceptions */
#define JB_OK                    0
#define JB_ERR_BAD_PARAM        -1   /* null pointer, non-positive size, bad enum            */
#define JB_ERR_BAD_QUANTIZATION -2   /* BadQuantizationError  pipeline/__init__.py:29,34,63  */
#define JB_ERR_EMPTY_ARRAY      -3   /* EmptyArrayError       util.py:30-31                  */
#define JB_ERR_UNSUPPORTED      -4   /* dct_size > JB_MAX_DCT_SIZE or block_size > JB_MAX_BLOCK_SIZE */
#define JB_ERR_WORKSPACE        -5   /* workspace smaller than jb_*_workspace_bytes()        */
#define JB_ERR_OUT_CAPACITY     -6   /* output buffer too small for the streams              */
#define JB_ERR_CUDA             -7   /* a CUDA runtime call failed                           */
#define JB_ERR_BAD_RLE_CODE     -8   /* BadRleCodeError       util.py:162-178 (size > 15)    */
#define JB_ERR_BAD_STREAM       -9   /* stream does not decode to the block count of the geometry */
#define JB_ERR_NO_DEVICE       -10   /* no CUDA device / library built without a usable arch */

#define JB_MAX_DCT_SIZE   32
#define JB_MAX_BLOCK_SIZE 255

/* Mirror of pipeline.Configuration (pipeline/__init__.py:50-64). */
typedef struct jb_params {
    int32_t height;      /* config.height : source rows                      */
    int32_t width;       /* config.width  : source columns                   */
    int32_t block_size;  /* --block_size  : sub-sampling box, >= 1           */
    int32_t dct_size;    /* --dct_size    : transform block, 1..32           */
    int32_t transform;   /* JB_TRANSFORM_*                                   */
    int32_t qmode;       /* JB_Q_*                                           */
    int32_t qparam;      /* --qkeep or --qdivisor, ignored otherwise         */
    int32_t flags;       /* JB_FLAG_* */
} jb_params;

#define JB_FLAG_FORCE_GENERIC 1  /* never take the specialised 8x8 / block_size 4 kernels */
#define JB_FLAG_NO_TMA        2  /* specialised kernels stage tiles with plain loads/stores */
#define JB_FLAG_NO_REFINE     4  /* skip the float64 re-evaluation of near-tie coefficients */
#define JB_FLAG_SERIAL_FRAMING 8 /* decoder: find block boundaries with the serial fallback walk only */
#define JB_FLAG_PDL 128          /* launch the kernels of the call with programmatic dependent launch: their prologues
                                   overlap the tail of the kernel before them on the stream (results unchanged) */
#define JB_FLAG_TILE_DECODER 64  /* decoder, 8x8 / block_size 4 kernel: store 32-row x 128-byte tiles (TMA tensor stores,
                                   or plain stores with JB_FLAG_NO_TMA) instead of whole chunk rows; kept for comparison */
#define JB_FLAG_REUSE_TABLES  16 /* the caller promises that this workspace was last used by a COMPLETED-OR-QUEUED call of
                                   the same direction with identical jb_params and n_planes (on the same stream, or
                                   ordered before this one) and has not been written since: the launches that build the
                                   tables and clean the control block are skipped -- a compress call is then three
                                   launches (fused transform kernel, scan, gather), a decompress call four (framing prep,
                                   walk, stitch, fused inverse kernel) */

/* Derived sizes (pipeline/run_length_encoding.py:80-88, pipeline/dct_padding.py:11-21). */
typedef struct jb_geometry {
    int32_t h1, w1;            /* after Padding + SubSampling                      */
    int32_t h2, w2;            /* after DCTPadding                                 */
    int32_t vb, hb;            /* blocks down / across                             */
    int32_t blocks_per_plane;  /* vb * hb                                          */
    int32_t max_block_bytes;   /* ceil((23 d^2 + 8) / 8): worst case of one block  */
    int32_t chunks_per_plane;  /* scheduling units: ceil(blocks_per_plane / 32), / 8 for dct_size >= 16 (informational:
                                * calls of at most 16384 large blocks are dealt in units of 2) */
    int32_t reserved;
} jb_geometry;

/* Status block written by the device (all entries uint64):
 *   [0] error code as a positive number (0 = ok, 8 = -JB_ERR_BAD_RLE_CODE, 9 = -JB_ERR_BAD_STREAM ...)
 *   [1] for BAD_RLE_CODE: the first bad code in stream order, packed as
 *         global block index (30 bits) << 34 | zigzag position (10) << 24 | run (4) << 20 |
 *         (amplitude + 2^19) (20 bits);   all ones when there is none
 *   [2] decoder: number of streams whose block boundaries were found by the serial fallback walk
 *   [3] reserved                                                                        */
#define JB_STATUS_WORDS 4

int jb_version(void);
const char* jb_strerror(int code);

/* Validate params and fill `geo` (host only, no CUDA call). */
int jb_geometry_of(const jb_params* p, jb_geometry* geo);

/* Worst-case total stream bytes for n_planes planes: blocks * max_block_bytes. */
size_t jb_max_stream_bytes(const jb_params* p, int n_planes);

size_t jb_compress_workspace_bytes(const jb_params* p, int n_planes);
size_t jb_decompress_workspace_bytes(const jb_params* p, int n_planes, size_t in_bytes);

/*
 * compress_band for a batch of planes: stages Padding, SubSampling, DCTPadding,
 * Normalization, BasisChange, Quantization, ZigzagOrder, RunLengthEncoding and
 * RleBytestream (the pipeline stage modules, step_index 0..8) fused in one pass.
 *
 *   d_planes      n_planes uint8 planes (see conventions)
 *   d_out         the streams, densely concatenated in plane order
 *   out_cap       capacity of d_out in bytes (jb_max_stream_bytes() always suffices)
 *   d_plane_off   n_planes + 1 uint64: stream p is d_out[off[p] .. off[p+1]); off[n_planes] = total
 *   d_status      JB_STATUS_WORDS uint64
 */
int jb_compress_planes(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                       const jb_params* p,
                       uint8_t* d_out, size_t out_cap, uint64_t* d_plane_off, uint64_t* d_status,
                       void* d_ws, size_t ws_bytes, void* stream);

/*
 * decompress_band for a batch of planes: the same nine stages inverted
 * (RleBytestream.invert ... Padding.invert) plus the uint8 cast of
 * Jpeg.decompress (pipeline/__init__.py:120-122).
 *
 *   d_in          stream bytes; stream p is d_in[d_plane_off[p] .. d_plane_off[p] + d_plane_len[p])
 *   in_bytes      host-known upper bound of max(off + len): sizes the launch
 *   d_plane_off / d_plane_len   n_planes uint64 each (device)
 *   d_planes_out  n_planes uint8 planes, height x width
 */
int jb_decompress_planes(const uint8_t* d_in, size_t in_bytes,
                         const uint64_t* d_plane_off, const uint64_t* d_plane_len, int n_planes,
                         const jb_params* p,
                         uint8_t* d_planes_out, size_t plane_stride, size_t row_pitch,
                         uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream);

/* Which block-boundary discovery a jb_decompress_planes call with these sizes runs (the container stores no
 * index, rle_byte_stream.py:74-88 reads sequentially; see csrc/jb_framing.cu): JB_FRAMING_STITCH = batches of
 * short streams, tile walks + one CTA per stream; JB_FRAMING_CHAIN = long streams (gigapixel planes), tile walks
 * chained over several kernels and CTAs per stream.
 * Host only; lets a test assert which path it exercised.  Negative: an error code. */
#define JB_FRAMING_STITCH 0
#define JB_FRAMING_CHAIN  1
int jb_decompress_framing_path(const jb_params* p, int n_planes, size_t in_bytes);

/* ---- stage-level entry points (parity hooks; same kernels, cut at the coefficient boundary) ---- */

/* Stages 0-6 + the int cast of RunLengthBlock.encode (run_length_encoding.py:16-17):
 * d_coeffs receives n_planes * blocks_per_plane * d*d int16 in zigzag order, blocks in
 * raster order (the (vb, hb, d^2) array of ZigzagOrder.execute, zigzag_order.py:85-99). */
int jb_stage_forward_coeffs(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                            const jb_params* p, int16_t* d_coeffs, uint64_t* d_status,
                            void* d_ws, size_t ws_bytes, void* stream);

/* Stages 7-8 (RunLengthEncoding.execute + RleBytestream.execute) from given int32
 * zigzag coefficients; only dct_size, the plane count and blocks_per_plane matter. */
int jb_stage_pack(const int32_t* d_coeffs, int n_planes, int blocks_per_plane, int dct_size,
                  uint8_t* d_out, size_t out_cap, uint64_t* d_plane_off, uint64_t* d_status,
                  void* d_ws, size_t ws_bytes, void* stream);
size_t jb_stage_pack_workspace_bytes(int n_planes, int blocks_per_plane, int dct_size);

/* Stages 8-7 inverted (RleBytestream.invert + RunLengthEncoding.invert): int16 zigzag
 * coefficients, n_planes * blocks_per_plane * d*d.  flags: JB_FLAG_SERIAL_FRAMING or 0. */
int jb_stage_unpack(const uint8_t* d_in, size_t in_bytes,
                    const uint64_t* d_plane_off, const uint64_t* d_plane_len, int n_planes,
                    int blocks_per_plane, int dct_size, int flags, int16_t* d_coeffs, uint64_t* d_status,
                    void* d_ws, size_t ws_bytes, void* stream);
size_t jb_stage_unpack_workspace_bytes(int n_planes, int blocks_per_plane, int dct_size, size_t in_bytes);

/* Stages 6-0 inverted from int16 zigzag coefficients to uint8 planes. */
int jb_stage_inverse_coeffs(const int16_t* d_coeffs, int n_planes, const jb_params* p,
                            uint8_t* d_planes_out, size_t plane_stride, size_t row_pitch,
                            uint64_t* d_status, void* d_ws, size_t ws_bytes, void* stream);

/* Float64 intermediates of the compress direction, as the reference's stages hand them on (pipeline/base.py:42-72;
 * each stage's execute() result is the next stage's input):
 *   JB_F64_SAMPLES      after Padding, SubSampling, DCTPadding (padding.py:8-12, subsampling.py:9-11,
 *                       dct_padding.py:8-9): n_planes x (vb*d) x (hb*d) box means;
 *   JB_F64_TRANSFORM    after BasisChange.execute (basis_change.py:11-26): n_planes x blocks_per_plane x d*d in
 *                       natural (u, v) order, blocks in raster order; the real part for the DFT;
 *   JB_F64_PREROUNDING  after the quantiser's scaling, the value np.round receives (quantizers.py:5-6, 16-17, 27-28,
 *                       47-49): same shape.
 * Computed operation for operation in the reference's order -- the arithmetic that decides rounding ties inside the
 * fused kernels -- so the arrays compare bit for bit (DCT; DFT with dct_size 8).  A parity hook, not a fast path.
 * Workspace: jb_compress_workspace_bytes(p, n_planes) is enough (only the tables are used). */
#define JB_F64_SAMPLES     0
#define JB_F64_TRANSFORM   1
#define JB_F64_PREROUNDING 2
int jb_stage_float64(const uint8_t* d_planes, size_t plane_stride, size_t row_pitch, int n_planes,
                     const jb_params* p, int which, double* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* ---- colour conversion on either side of the path -------------------------
 * Replaces PIL's im.convert('YCbCr') before Jpeg.compress reads the bands (compress.py:9,
 * pipeline/__init__.py:103-106) and im.convert('RGB') after Jpeg.decompress stacks them (decompress.py:10,
 * pipeline/__init__.py:120-124).  Bit-identical to Pillow on all 2^24 inputs (tables fitted against Pillow,
 * tools/derive_pil_tables.py).  d_rgb: interleaved uint8 [n_images][height][width][3] (rgb_pitch bytes per
 * row, image_stride bytes per image); d_planes: uint8 planes, image i owns planes 3i (Y), 3i+1 (Cb), 3i+2 (Cr)
 * -- exactly the plane order jb_compress_planes / jb_decompress_planes use for a batch of images. */
int jb_rgb_to_ycbcr_planes(const uint8_t* d_rgb, size_t image_stride, size_t rgb_pitch, int n_images,
                           int height, int width, uint8_t* d_planes, size_t plane_stride, size_t plane_pitch,
                           void* stream);
int jb_ycbcr_planes_to_rgb(const uint8_t* d_planes, size_t plane_stride, size_t plane_pitch, int n_images,
                           int height, int width, uint8_t* d_rgb, size_t image_stride, size_t rgb_pitch,
                           void* stream);

/* ---- containers of a batch of images -------------------------------------
 * Replaces the host loop over file_format.generate_data (file_format.py:86-93): container i = header, then for
 * Y, Cb, Cr a little-endian u32 length and the stream, taken from the output of jb_compress_planes on
 * 3 n planes (d_streams, d_plane_off[3 n + 1]).  `header` is a HOST pointer to file_format.create_header's
 * bytes (<= 256).  d_image_off receives n + 1 offsets into d_out; d_status[0] a positive error code. */
size_t jb_containers_max_bytes(int n_images, int header_len, size_t stream_bytes);
int jb_pack_containers(const uint8_t* d_streams, const uint64_t* d_plane_off, int n_images, const uint8_t* header,
                       int header_len, uint8_t* d_out, size_t out_cap, uint64_t* d_image_off, uint64_t* d_status,
                       void* stream);

/* ---- measurement hook (bench.py's roofline figure) ----------------------------------------------------------
 * Brackets the dominant kernel of the next calls with caller-owned CUDA events (cudaEvent_t handles, created
 * with timing enabled): jb_compress_planes records fwd_begin / fwd_end around its fused transform kernel only
 * (not the scan and gather launches that follow), jb_decompress_planes records inv_begin / inv_end around its
 * fused inverse kernel only (not the framing launches before it).  NULL handles switch a pair off; all NULL is
 * the default.  Process-wide and not thread safe: a profiling aid, not part of the data path; do not capture
 * calls into a CUDA graph while it is armed.  The reference has no counterpart. */
int jb_debug_kernel_events(void* fwd_begin, void* fwd_end, void* inv_begin, void* inv_end);

/* Kernels this library has launched in this process so far (every launch is counted, also while a CUDA graph is being
 * captured): the difference across a call -- or across the capture of a graph -- is the number of kernels it consists
 * of.  bench.py's gpu_launches comes from it. */
unsigned long long jb_debug_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JPEGB200_H */
