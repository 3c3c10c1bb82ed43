// orb_internal.cuh — shared declarations of liborb_b200 (sm_100a).  Not a public header.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/orb_b200.h"

#define ORB_MAX_LEVELS 12
#define ORB_EDGE 19        // EDGE_THRESHOLD   (reference ORBextractor.cc:72)
#define ORB_MINB 16        // EDGE_THRESHOLD-3 (reference ORBextractor.cc:801)
#define ORB_HALF_PATCH 15  // reference ORBextractor.cc:71
#define ORB_PATCH 31       // reference ORBextractor.cc:70
#define ORB_XOFF 32        // byte offset of interior pixel x = 0 inside a pyramid row (border occupies [13, 32))
#define ORB_NSTAGES 5      // pyramid, FAST cells, quadtree, blur, orientation+descriptors
#define ORB_PROF_RING 64
#define ORB_PROF_EVENTS 10
#define ORB_PIPE_SLOTS 8   // chunks in flight in the host-buffer pipeline
#define ORB_PIPE_CHUNK 64  // frames per chunk of the host-buffer pipeline

// thread-local error string ---------------------------------------------------------------------------
void orb_set_error(const char* fmt, ...);

#define ORB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            orb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return ORB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// per-level geometry, passed BY VALUE to kernels (lives in the constant bank) ---------------------------
struct LevelGeom {
    int w, h;                  // interior size of mvImagePyramid[l]
    int pitch, rows;           // bordered buffer: pitch bytes per row (multiple of 64), rows = h + 38
    int ioff;                  // byte offset of interior pixel (0,0) from the frame's level base: 19 * pitch + ORB_XOFF
    long long base;            // byte offset of frame 0 of this level in the pyramid arena (256-B aligned)
    long long frame_stride;    // pitch * rows
    int bpitch;                // blurred buffer pitch (w x h, no border)
    long long bbase, bframe_stride;
    int nCols, nRows, wCell, hCell, maxBX, maxBY;  // cell grid (reference ORBextractor.cc:801-817)
    int cell_base;             // number of cells in levels < l
    int quota;                 // mnFeaturesPerLevel[l]
    int nIni;                  // quadtree roots (reference ORBextractor.cc:566)
    float hX;
    int corner_cap;            // worst-case number of NMS survivors of this level
    long long corner_base;     // element offset (u64 records) of frame 0 / this level in the corner arena
    int node_cap;              // max quadtree list length (+ slack)
    int kp_base;               // offset of this level's slots in a frame's kept-keypoint array
    int xtab, ytab;            // offsets into the resize tables (entries of ResizeTap)
    int xwtab;                 // offset (entries of ResizeWord) of the packed per-output-word column table
    int fast_resize;           // 1: every output word's source taps fit 3 aligned source words (scale <= 4/3)
    int fast_G, fast_groups;   // FAST strips: cells per CTA, CTAs per cell row
    int fast_cta_base;         // number of FAST CTAs of levels < l (per frame)
    int border_words;          // words per bordered row from byte 12 on: left border | interior | right border (the pyramid kernels' row items)
    unsigned inv_wpr;          // ceil(2^32 / border_words)
    int nseg, seg_base, seg_threads, stage_bytes;   // staged pyramid kernel: segments of <= 256 bordered words per row, first ResizeSeg, CTA size, dynamic shared memory
    int blur_base, blur_wpr;   // blur: number of thread items of levels < l; words per row
    float scale;               // mvScaleFactor[l]
    float size;                // (float)(int)(PATCH_SIZE * scale)
};

struct Geometry {
    int nlevels, w, h;
    int ini_th, min_th;
    int one;                     // = 1: lets kernels build multipliers (1 << k) that the compiler cannot strength-reduce into ALU-pipe shifts
    int total_cells, total_kp_slots, max_node_cap, max_tile_bytes;
    int fast_rows;               // tile rows of the FAST strip kernel = the tallest cell sub-image of this geometry (<= 66): sizes its shared memory and the TMA box
    int fast_ctas, blur_items;   // per-frame grid sizes of the strip / blur kernels
    int l0_ni, l0_border_first;  // level-0 pass: 16-byte vectors wholly inside the interior per row; first item of the border vectors (a multiple of 32)
    unsigned l0_inv_ni, l0_inv_nb;   // ceil(2^32 / l0_ni), ceil(2^32 / (pitch / 16 - l0_ni))
    long long pyr_frame_total;   // not used for addressing (level-major layout), informational
    LevelGeom lv[ORB_MAX_LEVELS];
};

struct ResizeTap {  // one destination coordinate of resize(INTER_LINEAR): two source taps + Q11 weights
    unsigned short s0, s1;
    short c0, c1;
};
struct __align__(16) ResizeWord {  // 4 adjacent destination columns (one output word) of the fast resize path
    int wb;              // index of the first aligned source word
    unsigned sh0;        // 8 * (byte offset 0..3 of column 0's left tap inside word wb): the funnel shift that brings the
                         // 8 source bytes the word needs into two registers (A, B)
    unsigned sel01;      // PRMT selector over (A, B): (S[s], S[s+1]) of column 0 in bytes 0-1, of column 1 in bytes 2-3
    unsigned sel23;      // the same for columns 2 and 3
    unsigned cc[4];      // c0 | c1 << 16 per column (Q11)
};
struct ResizeSeg {   // one column segment of the staged pyramid kernel
    int w0;              // first staged source word of a row (a multiple of 4: 16-byte aligned)
    int nbytes;          // staged bytes per source row (a multiple of 16)
    int bw0, nw;         // first bordered word, number of words
};
// TMA descriptors of the pyramid levels (dims: row bytes, rows, frames of the arena) for the FAST strip loader
struct FastTmaps { CUtensorMap m[ORB_MAX_LEVELS]; };
#define ORB_TMA_BOX_W 256
#ifndef ORB_TAP_BOX_W
#define ORB_TAP_BOX_W 80      // descriptor tap window box: 80 bytes x 39 rows (columns px-19 .. px+19 from a 16-byte aligned start need 54;
                              // the TMA unit writes rows densely, and a 20-word row pitch spreads the rows over 8 bank offsets where 16 words
                              // give 2: orient_describe 0.477 -> 0.449 ms per 512 frames)
#endif
#define ORB_TAP_BOX_H 39

struct __align__(16) FastStrip {   // one CTA of fast_strip_kernel: up to fast_G consecutive valid cells of one cell row
    int level, i, j0, ncell;       // cell row, first cell column, number of valid cells
    int iniY, ch, X0, tw;          // cell sub-image rows [iniY, iniY+ch), tile columns [X0, X0+tw) (interior coords)
    int a, lw, nw, wlo;            // smem byte shift, words loaded per row, words with evaluated pixels, first such word
    unsigned inv_lw, inv_nw, inv_wc, inv_np;   // ceil(2^32 / x) magic numbers for the index divisions
    int np, w0p, pad0, pad1;       // stage A walks word PAIRS: pairs per row, first pair's first word (even, <= wlo)
};
#ifndef ORB_BLUR_ROWS
#define ORB_BLUR_ROWS 16   // output rows per blur thread
#endif
#ifndef ORB_RESIZE_ROWS
#define ORB_RESIZE_ROWS 16 // output rows per pyramid-kernel thread (throughput shape)
#endif

// 64-bit corner record: max() over records picks the reference's winner inside a quadtree node:
//   hi 32 bits = score << 24 | (0xFFFFFF - order)   order = (cell index << 12 | y_in_cell << 6 | x_in_cell)
//   lo 32 bits = x | y << 16     (x, y relative to the 16-px border == vToDistributeKeys coordinates)
__host__ __device__ inline unsigned long long corner_pack(int x, int y, int score, int order) {
    return ((unsigned long long)(((unsigned)score << 24) | (0xFFFFFFu - (unsigned)order)) << 32) |
           (unsigned)(x | (y << 16));
}
__host__ __device__ inline int corner_x(unsigned long long r) { return (int)(r & 0xFFFF); }
__host__ __device__ inline int corner_y(unsigned long long r) { return (int)((r >> 16) & 0xFFFF); }
__host__ __device__ inline int corner_score(unsigned long long r) { return (int)(r >> 56); }
__host__ __device__ inline int corner_order(unsigned long long r) { return (int)(0xFFFFFFu - ((r >> 32) & 0xFFFFFFu)); }

struct orb_ctx {
    int device = 0;
    int nfeatures = 0, nlevels = 0, ini_th = 0, min_th = 0;
    double scale_factor = 0;
    float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS], sigma2[ORB_MAX_LEVELS], inv_sigma2[ORB_MAX_LEVELS];
    int quota[ORB_MAX_LEVELS];
    int max_batch = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    long long launches = 0;
    // geometry-dependent state (rebuilt when the image size changes)
    bool have_geom = false;
    Geometry g;
    Geometry gl;               // per-launch copy of g with the bases shifted to the chunk's first frame
    // host-API pipeline (orb_extract_batch): copy streams + per-chunk events
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr, st_c2 = nullptr;
    cudaStream_t st_aux[2] = {nullptr, nullptr};   // border -> blur chain next to FAST -> quadtree (one per compute stream)
    cudaEvent_t ev_pyr[2] = {}, ev_blur[2] = {};
    cudaEvent_t ev_head = nullptr, ev_early = nullptr;   // latency path: FAST + quadtree of the lower levels run next to the pyramid's tail launch
    bool overlap = false;
    cudaEvent_t ev_in[ORB_PIPE_SLOTS] = {}, ev_done[ORB_PIPE_SLOTS] = {}, ev_out[ORB_PIPE_SLOTS] = {};
    int last_frames = 0;
    uint8_t* d_in = nullptr;  size_t in_bytes = 0;      // staging of host input frames
    uint8_t* d_pyr = nullptr; size_t pyr_bytes = 0;     // bordered pyramids, level-major
    uint8_t* d_blur = nullptr; size_t blur_bytes = 0;   // blurred levels
    unsigned long long* d_corners = nullptr; size_t corner_elems = 0;
    unsigned short* d_node_of_key = nullptr;
    int* d_corner_count = nullptr;                      // [2][max_batch][nlevels]: corner counts, tie-at-cut counts
    unsigned long long* d_kept = nullptr;               // [max_batch][total_kp_slots]
    int* d_kept_count = nullptr;                        // [max_batch][nlevels]
    ResizeTap* d_taps = nullptr;
    ResizeWord* d_wtaps = nullptr;
    ResizeSeg* d_rsegs = nullptr;
    FastStrip* d_strips = nullptr;
    // FAST variant choice (orb_fast.cu): 0 / 1 = stage A evaluates iniThFAST only / both thresholds, 2 = by the share of empty cells
    // the kernel reports (ORB_B200_FAST_DUAL; default 2)
    int fast_dual_mode = 2;
    bool fast_dual_now = false, fast_stats_pending = false;
    unsigned fast_launch_no = 0;
    int* d_fast_stats = nullptr;      // 64 x (empty cells, cells)
    int* h_fast_stats = nullptr;      // pinned copy
    cudaEvent_t ev_fast_stats = nullptr;
    FastTmaps tmaps;              // valid when use_tma
    FastTmaps* d_tmaps = nullptr; // device copy (the TMA unit reads the descriptor from global memory)
    FastTmaps* d_btmaps = nullptr; // the same for the BLURRED levels (dims: row bytes, rows, frames): descriptor tap windows (orient_describe_kernel)
    bool use_tma = false;
    uint2* d_mom_tab = nullptr;   // IC_Angle weight table [4 alignments][288 items] (orient_describe_kernel)
    bool fast_attr_set = false, qt_attr_set = false;
    // single-frame latency path: the whole call (H2D from the pinned staging buffer, 13 kernels with the border -> blur
    // chain next to FAST -> quadtree, D2H into the pinned mirrors) as ONE CUDA graph, re-captured when a key field changes
    int use_graph = -1;                    // -1: not decided yet (ORB_B200_GRAPH=0 disables)
    cudaGraphExec_t graph_exec = nullptr;
    struct GraphKey { int fmt, w, h, rows, ocap; const void *d_in, *h_in, *h_kps, *h_desc, *d_kps; } graph_key = {};
    int graph_warm_w = 0, graph_warm_h = 0, graph_warm_fmt = -1;   // geometry / format that already ran once without a graph
    long long graph_launches = 0;          // kernels inside the captured graph
    orb_kp* d_kps_out = nullptr; uint8_t* d_desc_out = nullptr; int* d_n_out = nullptr; int out_cap = 0;
    orb_kp* h_kps = nullptr; uint8_t* h_desc = nullptr; int* h_n = nullptr; uint8_t* h_in = nullptr;  // pinned
    size_t h_in_bytes = 0;
    int h_out_cap = 0;
    uint8_t *d_scratch = nullptr, *h_scratch = nullptr;   // grow-only slabs of orb_stereo_match
    size_t d_scratch_cap = 0, h_scratch_cap = 0;
    // per-stage CUDA-event timers (orb_profile_enable / orb_profile_read): a ring of event sets so that reading
    // never stalls the stream; stage s of a call = elapsed(ev[s], ev[s+1])
    bool profile = false;
    cudaEvent_t prof_ev[ORB_PROF_RING][ORB_PROF_EVENTS] = {};   // main stream 0..5, [6..8] border/blur chain, [9] FAST start
    bool prof_pending[ORB_PROF_RING] = {};
    int prof_frames[ORB_PROF_RING] = {};
    int prof_head = 0;
    double prof_ms[ORB_NSTAGES] = {};
    long long prof_calls = 0, prof_total_frames = 0;
};
int orb_profile_harvest(orb_ctx* c, int slot);

// kernels' launchers (orb_extract_kernels.cu)
int orb_launch_pyramid(orb_ctx* c, const Geometry& g, const uint8_t* d_imgs, int pixel_format, int nframes, size_t row_stride,
                       size_t frame_stride, cudaStream_t st, int phase = 0, int* tail_first_out = nullptr);
inline int orb_pix_channels(int fmt) { return fmt == ORB_PIX_GRAY8 ? 1 : (fmt == ORB_PIX_BGR8 || fmt == ORB_PIX_RGB8) ? 3 : 4; }
void orb_carveout_pyramid(int pct);
void orb_carveout_blur(int pct);
void orb_carveout_fast(int pct);
void orb_carveout_extract(int pct);
int orb_launch_fast(orb_ctx* c, const Geometry& g, int* d_corner_count, int nframes, int f0, cudaStream_t st, int level_begin = 0,
                    int level_end = ORB_MAX_LEVELS);
int orb_launch_blur(orb_ctx* c, const Geometry& g, int nframes, cudaStream_t st);
// frames [f0, f0 + nframes) of the arena; all pointers address the chunk's first frame; asynchronous on st
int orb_launch_extract(orb_ctx* c, const uint8_t* d_imgs, int pixel_format, int nframes, int f0, size_t row_stride,
                       size_t frame_stride, orb_kp* d_kps, uint8_t* d_desc, int cap, int* d_n_out, cudaStream_t st);
