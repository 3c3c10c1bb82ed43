// orb_extract_host.cu — host side of the extractor C ABI (include/orb_b200.h): context, tables, geometry,
// device arena and the orb_extract* entry points.  Mirrors the reference constructor
// (orb_slam2/src/ORBextractor.cc:416-479) and ComputePyramid's size rules (:1152-1165).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "orb_internal.cuh"

static thread_local char g_err[512] = "";

void orb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* orb_last_error(void) { return g_err; }
extern "C" int orb_version(void) { return 100; }
extern "C" int orb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

namespace {

inline int cv_round_f(float v) { return (int)lrintf(v); }  // cvRound: half-to-even
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void free_geometry_buffers(orb_ctx* c) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    c->graph_warm_w = c->graph_warm_h = 0; c->graph_warm_fmt = -1;
    cudaFree(c->d_in); cudaFree(c->d_pyr); cudaFree(c->d_blur); cudaFree(c->d_corners); cudaFree(c->d_node_of_key);
    cudaFree(c->d_corner_count); cudaFree(c->d_kept); cudaFree(c->d_kept_count); cudaFree(c->d_taps);
    cudaFree(c->d_wtaps); cudaFree(c->d_rsegs); cudaFree(c->d_strips); cudaFree(c->d_tmaps); cudaFree(c->d_btmaps); cudaFree(c->d_kps_out); cudaFree(c->d_desc_out); cudaFree(c->d_n_out);
    cudaFreeHost(c->h_kps); cudaFreeHost(c->h_desc); cudaFreeHost(c->h_n); cudaFreeHost(c->h_in);
    c->h_out_cap = 0;
    c->d_in = c->d_pyr = c->d_blur = nullptr; c->d_corners = nullptr; c->d_node_of_key = nullptr;
    c->d_corner_count = c->d_kept_count = nullptr; c->d_kept = nullptr; c->d_taps = nullptr;
    c->d_wtaps = nullptr; c->d_rsegs = nullptr; c->d_strips = nullptr; c->d_tmaps = nullptr; c->d_btmaps = nullptr; c->d_kps_out = nullptr; c->d_desc_out = nullptr; c->d_n_out = nullptr;
    c->h_kps = nullptr; c->h_desc = nullptr; c->h_n = nullptr; c->h_in = nullptr;
    c->in_bytes = c->h_in_bytes = 0; c->out_cap = 0;
    c->have_geom = false;
}

// resize(INTER_LINEAR) tap tables of one axis (OpenCV resize.cpp, 8u fixed-point path)
void make_taps(int sn, int dn, bool vertical, ResizeTap* out) {
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        int s0, s1;
        if (!vertical) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= sn - 1) { s = sn - 1; f = 0.f; }
            s0 = s; s1 = std::min(s + 1, sn - 1);
        } else {  // rows are clipped, the weights are kept
            s0 = std::min(std::max(s, 0), sn - 1);
            s1 = std::min(std::max(s + 1, 0), sn - 1);
        }
        out[d].s0 = (unsigned short)s0; out[d].s1 = (unsigned short)s1;
        out[d].c0 = (short)cv_round_f((1.f - f) * 2048.f);
        out[d].c1 = (short)cv_round_f(f * 2048.f);
    }
}

// (re)build everything that depends on the image size
int build_geometry(orb_ctx* c, int w, int h) {
    if (c->have_geom && c->g.w == w && c->g.h == h) return ORB_OK;
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    free_geometry_buffers(c);
    Geometry& g = c->g;
    memset(&g, 0, sizeof(g));
    g.nlevels = c->nlevels; g.w = w; g.h = h; g.ini_th = c->ini_th; g.min_th = c->min_th; g.one = 1;
    const int F = c->max_batch;
    long long pyr_off = 0, blur_off = 0, corner_off = 0;
    int cells = 0, kp_slots = 0, taps = 0, wtaps = 0, max_node_cap = 0, fast_ctas = 0, blur_items = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom& L = g.lv[l];
        L.w = cv_round_f((float)w * c->inv_scale[l]);  // ORBextractor.cc:1159
        L.h = cv_round_f((float)h * c->inv_scale[l]);
        if (L.w > 4096 + 32 || L.h > 4096 + 32) { orb_set_error("image larger than 4128 px is not supported"); return ORB_ERR_INVALID; }
        // row layout: [0,13) dead | [13,32) left border | [32, 32+w) interior (16-B aligned) | 19 right border | pad
        L.pitch = (int)round_up(ORB_XOFF + L.w + ORB_EDGE, 64);
        L.rows = L.h + 2 * ORB_EDGE;
        L.ioff = ORB_EDGE * L.pitch + ORB_XOFF;
        L.frame_stride = (long long)L.pitch * L.rows;
        L.base = pyr_off; pyr_off += L.frame_stride * F;
        L.bpitch = (int)round_up(L.w, 64);
        L.bframe_stride = (long long)L.bpitch * L.h;
        L.bbase = blur_off; blur_off += L.bframe_stride * F;
        // cell grid, ORBextractor.cc:801-817 (float arithmetic as in the reference)
        L.maxBX = L.w - ORB_EDGE + 3; L.maxBY = L.h - ORB_EDGE + 3;
        const float width = (float)(L.maxBX - ORB_MINB), height = (float)(L.maxBY - ORB_MINB);
        L.nCols = (int)(width / 30.f); L.nRows = (int)(height / 30.f);
        if (L.nCols <= 0 || L.nRows <= 0) {
            orb_set_error("level %d (%dx%d) is smaller than one 30-px cell: the reference divides by zero here", l, L.w, L.h);
            return ORB_ERR_TOO_SMALL;
        }
        L.wCell = (int)ceilf(width / L.nCols); L.hCell = (int)ceilf(height / L.nRows);
        L.cell_base = cells; cells += L.nCols * L.nRows;
        if (L.nCols * L.nRows > 4096) { orb_set_error("more than 4096 cells per level is not supported"); return ORB_ERR_INVALID; }
        if (L.wCell > 60 || L.hCell > 60) { orb_set_error("internal: cell larger than 60 px"); return ORB_ERR_INVALID; }
        L.quota = c->quota[l];
        // quadtree roots, ORBextractor.cc:566-568
        L.nIni = (int)roundf((float)(L.maxBX - ORB_MINB) / (L.maxBY - ORB_MINB));
        if (L.nIni <= 0) { orb_set_error("level %d is taller than 2x its width: the reference divides by zero here", l); return ORB_ERR_TOO_SMALL; }
        L.hX = (float)(L.maxBX - ORB_MINB) / L.nIni;
        // worst case of NMS survivors: one per 2x2 block of every cell's evaluated area
        long long cc = 0;
        for (int i = 0; i < L.nRows; ++i) {
            const int iniY = ORB_MINB + i * L.hCell;
            if (iniY >= L.maxBY - 3) continue;
            const int ch = std::min(iniY + L.hCell + 6, L.maxBY) - iniY;
            for (int j = 0; j < L.nCols; ++j) {
                const int iniX = ORB_MINB + j * L.wCell;
                if (iniX >= L.maxBX - 6) continue;
                const int cw = std::min(iniX + L.wCell + 6, L.maxBX) - iniX;
                if (cw < 7 || ch < 7) continue;
                cc += (long long)((cw - 6 + 1) / 2) * ((ch - 6 + 1) / 2);
            }
        }
        L.corner_cap = (int)round_up((size_t)std::max<long long>(cc, 8), 8);
        L.corner_base = corner_off; corner_off += (long long)L.corner_cap * F;
        L.node_cap = (int)round_up((size_t)std::max(L.quota + 3, 4 * L.nIni) + L.nIni + 8, 8);
        max_node_cap = std::max(max_node_cap, L.node_cap);
        L.kp_base = kp_slots; kp_slots += L.node_cap;
        L.xtab = taps; taps += L.w;
        L.ytab = taps; taps += L.h;
        L.xwtab = wtaps; wtaps += (ORB_XOFF + L.w + ORB_EDGE + 3) / 4 - (ORB_XOFF - ORB_EDGE) / 4;   // one entry per bordered word
        L.scale = c->scale[l];
        L.size = (float)(int)(ORB_PATCH * c->scale[l]);  // ORBextractor.cc:874
        // FAST strips: one CTA = fast_G consecutive cells of one cell row (tile width fast_G * wCell + 6 <= 256)
        const int gmax = std::max(1, 231 / L.wCell);   // 19 (alignment shift) + tile <= 256 bytes = one TMA box row
        L.fast_groups = (L.nCols + gmax - 1) / gmax;
        L.fast_G = (L.nCols + L.fast_groups - 1) / L.fast_groups;
        L.fast_cta_base = fast_ctas; fast_ctas += L.nRows * L.fast_groups;
        // pyramid row items: a bordered row is written as the words from byte 12 (column -20, a dead byte) to the end of the
        // right border; level 0 is written as 16-byte vectors of the whole pitch
        const int first_w = (ORB_XOFF - ORB_EDGE) / 4;                          // word holding byte 13
        const int end_w = (ORB_XOFF + L.w + ORB_EDGE + 3) / 4;                   // one past the last border word
        L.border_words = end_w - first_w;
        L.inv_wpr = 0xFFFFFFFFu / (unsigned)L.border_words + 1u;
        // blur: one thread = one output word x ORB_BLUR_ROWS rows
        L.blur_wpr = (L.w + 3) / 4;
        L.blur_base = blur_items;   // a multiple of the CTA size: every CTA of the blur kernel lies inside ONE level (uniform level data)
        blur_items += (L.blur_wpr * ((L.h + ORB_BLUR_ROWS - 1) / ORB_BLUR_ROWS) + 255) / 256 * 256;
    }
    if (max_node_cap > 65535) { orb_set_error("nfeatures too large"); return ORB_ERR_INVALID; }
    if ((size_t)max_node_cap * 80 > 200 * 1024) { orb_set_error("nfeatures too large for the quadtree kernel"); return ORB_ERR_INVALID; }
    g.total_cells = cells; g.total_kp_slots = kp_slots; g.max_node_cap = max_node_cap;
    g.blur_items = blur_items;
    g.l0_ni = g.lv[0].w >> 4;   // level-0 pass: interior vectors [2, 2 + l0_ni) of a bordered row are plain copies
    g.l0_border_first = (g.l0_ni * g.lv[0].h + 31) / 32 * 32;
    g.l0_inv_ni = 0xFFFFFFFFu / (unsigned)std::max(g.l0_ni, 1) + 1u;
    g.l0_inv_nb = 0xFFFFFFFFu / (unsigned)((g.lv[0].pitch >> 4) - g.l0_ni) + 1u;
    g.pyr_frame_total = pyr_off / F;
    // resize taps: per-column / per-row tables + the packed per-output-word table of the fast path
    std::vector<ResizeTap> h_taps(std::max(taps, 1));
    std::vector<ResizeWord> h_wtaps(std::max(wtaps, 1));
    std::vector<ResizeSeg> h_rsegs;
    for (int l = 1; l < g.nlevels; ++l) {
        LevelGeom& L = g.lv[l];
        ResizeTap* xt = h_taps.data() + L.xtab;
        make_taps(g.lv[l - 1].w, L.w, false, xt);
        make_taps(g.lv[l - 1].h, L.h, true, h_taps.data() + L.ytab);
        L.fast_resize = 1;
        // one entry per word of the bordered row: the 4 columns of word bw are x0 .. x0 + 3, x0 = 4 * bw - 20; a border column
        // takes the taps of the column it reflects (copyMakeBorder(REFLECT_101) of the resized level).  All 8 tap bytes must
        // lie in the 8 bytes that start at the leftmost tap (true for scale factors <= 4/3, reflected words included).
        for (int bw = 0; bw < L.border_words; ++bw) {
            ResizeWord& rw = h_wtaps[L.xwtab + bw];
            memset(&rw, 0, sizeof(rw));
            const int x0 = 4 * bw - (ORB_XOFF - 12);
            int col[4], smin = 1 << 30;
            for (int p = 0; p < 4; ++p) {
                int x = x0 + p;
                x = x < 0 ? -x : x;
                if (x >= L.w) x = 2 * L.w - 2 - x;
                col[p] = std::min(std::max(x, 0), L.w - 1);    // bytes beyond the 19-px border (row padding) are dead
                smin = std::min(smin, (int)xt[col[p]].s0);
            }
            rw.wb = smin >> 2;
            const int off0 = smin - 4 * rw.wb;                 // 0..3
            rw.sh0 = 8u * (unsigned)off0;
            unsigned sel[4] = {0, 0, 0, 0};
            for (int p = 0; p < 4; ++p) {
                const ResizeTap& t = xt[col[p]];
                const int rel = (int)t.s0 - smin;              // left tap relative to the leftmost one: both taps must lie in 8 bytes
                if (rel < 0 || rel > 6) L.fast_resize = 0;
                sel[p] = (unsigned)(rel & 7) | ((unsigned)((rel + 1) & 7) << 4);
                rw.cc[p] = (unsigned)(unsigned short)t.c0 | ((unsigned)(unsigned short)t.c1 << 16);
            }
            rw.sel01 = sel[0] | (sel[1] << 8);
            rw.sel23 = sel[2] | (sel[3] << 8);
        }
        // staged kernel: column segments of <= 256 bordered words; a segment stages the source words [w0, w0 + nbytes / 4) of
        // every source row its strip touches
        L.nseg = (L.border_words + 255) / 256;
        L.seg_threads = (int)round_up((size_t)(L.border_words + L.nseg - 1) / L.nseg, 32);
        L.seg_base = (int)h_rsegs.size();
        int max_bytes = 16;
        for (int sgi = 0; sgi < L.nseg; ++sgi) {
            ResizeSeg sg;
            sg.bw0 = sgi * L.seg_threads;
            sg.nw = std::min(L.seg_threads, L.border_words - sg.bw0);
            int wmin = 1 << 30, wmax = 0;
            for (int bw = sg.bw0; bw < sg.bw0 + sg.nw; ++bw) {
                wmin = std::min(wmin, h_wtaps[L.xwtab + bw].wb);
                wmax = std::max(wmax, h_wtaps[L.xwtab + bw].wb + 2);
            }
            sg.w0 = wmin & ~3;
            sg.nbytes = (int)round_up((size_t)(wmax + 1 - sg.w0) * 4, 16);
            max_bytes = std::max(max_bytes, sg.nbytes);
            h_rsegs.push_back(sg);
        }
        int max_rows = 2;
        const ResizeTap* yt = h_taps.data() + L.ytab;
        for (int y0 = 0; y0 < L.h; y0 += ORB_RESIZE_ROWS)
            max_rows = std::max(max_rows, (int)yt[std::min(y0 + ORB_RESIZE_ROWS, L.h) - 1].s1 - (int)yt[y0].s0 + 1);
        L.stage_bytes = max_rows * max_bytes;
        if (L.stage_bytes > 200 * 1024) L.fast_resize = 0;
    }

    // FAST strips: valid cells only (ORBextractor.cc:822-837 skip rules), everything a CTA needs precomputed
    std::vector<FastStrip> strips;
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom& L = g.lv[l];
        L.fast_cta_base = (int)strips.size();
        for (int i = 0; i < L.nRows; ++i) {
            const int iniY = ORB_MINB + i * L.hCell;
            if (iniY >= L.maxBY - 3) continue;
            const int ch = std::min(iniY + L.hCell + 6, L.maxBY) - iniY;
            if (ch < 7) continue;                                   // cv::FAST returns nothing on such a sub-image
            for (int j0 = 0; j0 < L.nCols; j0 += L.fast_G) {
                int ncell = 0, X1 = 0;
                for (int j = j0; j < std::min(j0 + L.fast_G, L.nCols); ++j) {
                    const int iniX = ORB_MINB + j * L.wCell;
                    if (iniX >= L.maxBX - 6) break;
                    const int maxX = std::min(iniX + L.wCell + 6, L.maxBX);
                    if (maxX - iniX < 7) break;
                    ncell = j - j0 + 1; X1 = maxX;
                }
                if (ncell == 0) continue;
                FastStrip s;
                memset(&s, 0, sizeof(s));
                s.level = l; s.i = i; s.j0 = j0; s.ncell = ncell; s.iniY = iniY; s.ch = ch;
                s.X0 = ORB_MINB + j0 * L.wCell; s.tw = X1 - s.X0;
                // the tile starts at a 16-byte aligned global address (a TMA requirement; interior rows are 16-byte aligned,
                // so that is X0 rounded down to 16) with at least one spare word on the left: shift a in [4, 19]
                s.a = (s.X0 & 15) < 4 ? (s.X0 & 15) + 16 : (s.X0 & 15);
                s.lw = (s.a + s.tw + 3) >> 2;
                const int sb_lo = s.a + 3, sb_hi = s.a + 3 + (s.tw - 6);
                s.wlo = sb_lo >> 2; s.nw = ((sb_hi - 1) >> 2) - s.wlo + 1;
                s.inv_lw = 0xFFFFFFFFu / (unsigned)s.lw + 1u;
                s.inv_nw = 0xFFFFFFFFu / (unsigned)s.nw + 1u;
                s.inv_wc = 0xFFFFFFFFu / (unsigned)L.wCell + 1u;
                s.w0p = s.wlo & ~1; s.np = std::max(2, ((s.wlo + s.nw - 1) >> 1) - (s.w0p >> 1) + 1);   // >= 2: the 32-bit magic number of a division by 1 does not exist (a surplus pair has no valid pixel)
                s.inv_np = 0xFFFFFFFFu / (unsigned)s.np + 1u;
                strips.push_back(s);
            }
        }
    }
    g.fast_ctas = (int)strips.size();
    g.fast_rows = 7;
    for (const FastStrip& s : strips) g.fast_rows = std::max(g.fast_rows, s.ch);
    if (strips.empty()) strips.push_back(FastStrip());
    ORB_CUDA(cudaMalloc(&c->d_strips, sizeof(FastStrip) * strips.size()));
    ORB_CUDA(cudaMemcpyAsync(c->d_strips, strips.data(), sizeof(FastStrip) * strips.size(), cudaMemcpyHostToDevice, c->stream));
    if (!c->d_fast_stats) {   // once per context
        ORB_CUDA(cudaMalloc(&c->d_fast_stats, 128 * sizeof(int)));
        ORB_CUDA(cudaMemsetAsync(c->d_fast_stats, 0, 128 * sizeof(int), c->stream));
        ORB_CUDA(cudaHostAlloc(&c->h_fast_stats, 128 * sizeof(int), cudaHostAllocDefault));
        ORB_CUDA(cudaEventCreateWithFlags(&c->ev_fast_stats, cudaEventDisableTiming));
        const char* e = getenv("ORB_B200_FAST_DUAL");
        c->fast_dual_mode = (e && (e[0] == '0' || e[0] == '1') && !e[1]) ? e[0] - '0' : 2;
    }

    c->pyr_bytes = (size_t)pyr_off; c->blur_bytes = (size_t)blur_off; c->corner_elems = (size_t)corner_off;
    ORB_CUDA(cudaMalloc(&c->d_pyr, c->pyr_bytes + 256));
    ORB_CUDA(cudaMalloc(&c->d_blur, c->blur_bytes + 256));
    ORB_CUDA(cudaMalloc(&c->d_corners, c->corner_elems * sizeof(unsigned long long)));
    ORB_CUDA(cudaMalloc(&c->d_node_of_key, c->corner_elems * sizeof(unsigned short)));
    ORB_CUDA(cudaMalloc(&c->d_corner_count, sizeof(int) * 2 * (size_t)F * g.nlevels));
    ORB_CUDA(cudaMalloc(&c->d_kept, sizeof(unsigned long long) * (size_t)F * kp_slots));
    ORB_CUDA(cudaMalloc(&c->d_kept_count, sizeof(int) * (size_t)F * g.nlevels));
    ORB_CUDA(cudaMalloc(&c->d_taps, sizeof(ResizeTap) * h_taps.size()));
    ORB_CUDA(cudaMalloc(&c->d_wtaps, sizeof(ResizeWord) * h_wtaps.size()));
    ORB_CUDA(cudaMemcpyAsync(c->d_taps, h_taps.data(), sizeof(ResizeTap) * h_taps.size(), cudaMemcpyHostToDevice, c->stream));
    ORB_CUDA(cudaMemcpyAsync(c->d_wtaps, h_wtaps.data(), sizeof(ResizeWord) * h_wtaps.size(), cudaMemcpyHostToDevice, c->stream));
    if (h_rsegs.empty()) h_rsegs.push_back(ResizeSeg());
    ORB_CUDA(cudaMalloc(&c->d_rsegs, sizeof(ResizeSeg) * h_rsegs.size()));
    ORB_CUDA(cudaMemcpyAsync(c->d_rsegs, h_rsegs.data(), sizeof(ResizeSeg) * h_rsegs.size(), cudaMemcpyHostToDevice, c->stream));
    // TMA descriptors for the FAST strip loader (cuTensorMapEncodeTiled through the runtime's driver entry point: no
    // link dependency on libcuda).  Without them the kernel falls back to its LDG -> STS loader.
    c->use_tma = false;
    {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (!getenv("ORB_B200_NO_TMA") &&
            cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            bool ok = true;
            for (int l = 0; l < g.nlevels && ok; ++l) {
                const LevelGeom& L = g.lv[l];
                const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)F};
                const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.frame_stride};   // bytes, dims 1 and 2
                const cuuint32_t box[3] = {ORB_TMA_BOX_W, (cuuint32_t)g.fast_rows, 1};   // one box = the tallest strip of this geometry
                const cuuint32_t estr[3] = {1, 1, 1};
                const CUresult r = ((EncodeFn)fn)(&c->tmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d_pyr + L.base, dims, strides, box,
                                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                ok = (r == CUDA_SUCCESS);
            }
            if (ok) {
                ORB_CUDA(cudaMalloc(&c->d_tmaps, sizeof(FastTmaps)));
                ORB_CUDA(cudaMemcpyAsync(c->d_tmaps, &c->tmaps, sizeof(FastTmaps), cudaMemcpyHostToDevice, c->stream));
            }
            // the blurred levels, for the 64 x 39-byte tap window of one keypoint (orient_describe_kernel): rows / bytes past the
            // level are zero-filled by the TMA unit and never addressed by a tap
            FastTmaps bt;
            for (int l = 0; l < g.nlevels && ok; ++l) {
                const LevelGeom& L = g.lv[l];
                const cuuint64_t dims[3] = {(cuuint64_t)L.bpitch, (cuuint64_t)L.h, (cuuint64_t)F};
                const cuuint64_t strides[2] = {(cuuint64_t)L.bpitch, (cuuint64_t)L.bframe_stride};
                const cuuint32_t box[3] = {ORB_TAP_BOX_W, ORB_TAP_BOX_H, 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                const CUresult r = ((EncodeFn)fn)(&bt.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d_blur + L.bbase, dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                ok = (r == CUDA_SUCCESS);
            }
            if (ok && !getenv("ORB_B200_NO_TAP_TMA")) {
                ORB_CUDA(cudaMalloc(&c->d_btmaps, sizeof(FastTmaps)));
                ORB_CUDA(cudaMemcpyAsync(c->d_btmaps, &bt, sizeof(FastTmaps), cudaMemcpyHostToDevice, c->stream));
                ORB_CUDA(cudaStreamSynchronize(c->stream));   // bt is a local
            }
            c->use_tma = ok;
        } else {
            cudaGetLastError();
        }
    }
    ORB_CUDA(cudaMemsetAsync(c->d_pyr, 0, c->pyr_bytes + 256, c->stream));
    ORB_CUDA(cudaMemsetAsync(c->d_blur, 0, c->blur_bytes + 256, c->stream));
    ORB_CUDA(cudaMemsetAsync(c->d_kept_count, 0, sizeof(int) * (size_t)F * g.nlevels, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_geom = true;
    c->last_frames = 0;
    return ORB_OK;
}

// device output rows [max_batch][dcap] (+ pinned mirrors, used when the caller's buffers are pageable)
int ensure_outputs(orb_ctx* c, int dcap, bool need_host_mirror) {
    const int F = c->max_batch;
    if (c->out_cap < dcap) {
        ORB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_kps_out); cudaFree(c->d_desc_out); cudaFree(c->d_n_out);
        cudaFreeHost(c->h_kps); cudaFreeHost(c->h_desc); cudaFreeHost(c->h_n);
        c->d_kps_out = nullptr; c->d_desc_out = nullptr; c->d_n_out = nullptr;
        c->h_kps = nullptr; c->h_desc = nullptr; c->h_n = nullptr; c->h_out_cap = 0;
        ORB_CUDA(cudaMalloc(&c->d_kps_out, sizeof(orb_kp) * (size_t)F * dcap));
        ORB_CUDA(cudaMalloc(&c->d_desc_out, (size_t)32 * F * dcap));
        ORB_CUDA(cudaMalloc(&c->d_n_out, sizeof(int) * F));
        ORB_CUDA(cudaMallocHost(&c->h_n, sizeof(int) * F));
        c->out_cap = dcap;
    }
    if (need_host_mirror && c->h_out_cap < c->out_cap) {
        cudaFreeHost(c->h_kps); cudaFreeHost(c->h_desc);
        c->h_kps = nullptr; c->h_desc = nullptr;
        ORB_CUDA(cudaMallocHost(&c->h_kps, sizeof(orb_kp) * (size_t)F * c->out_cap));
        ORB_CUDA(cudaMallocHost(&c->h_desc, (size_t)32 * F * c->out_cap));
        c->h_out_cap = c->out_cap;
    }
    return ORB_OK;
}

int ensure_input_staging(orb_ctx* c, bool need_host_mirror, int channels) {
    const size_t in_bytes = (size_t)c->max_batch * c->g.w * c->g.h * channels;
    if (c->in_bytes < in_bytes) {
        ORB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_in); c->d_in = nullptr;
        ORB_CUDA(cudaMalloc(&c->d_in, in_bytes));
        c->in_bytes = in_bytes;
    }
    if (need_host_mirror && c->h_in_bytes < in_bytes) {
        cudaFreeHost(c->h_in); c->h_in = nullptr;
        ORB_CUDA(cudaMallocHost(&c->h_in, in_bytes));
        c->h_in_bytes = in_bytes;
    }
    return ORB_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

int check_device() {
    if (orb_device_count() <= 0) {
        orb_set_error("no CUDA device visible: liborb_b200 has no CPU fallback");
        return ORB_ERR_NO_DEVICE;
    }
    return ORB_OK;
}

}  // namespace

int orb_profile_harvest(orb_ctx* c, int slot) {
    cudaEvent_t* ev = c->prof_ev[slot];
    ORB_CUDA(cudaEventSynchronize(ev[5]));
    ORB_CUDA(cudaEventSynchronize(ev[8]));
    // stage = {pyramid (interior + borders), FAST, quadtree, blur, orientation + descriptors}; the border -> blur chain
    // [6..8] may have run on a second stream next to FAST -> quadtree [9, 2, 3]
    const int from[ORB_NSTAGES][2] = {{0, 6}, {9, -1}, {2, -1}, {7, -1}, {4, -1}};
    const int to[ORB_NSTAGES][2] = {{1, 7}, {2, -1}, {3, -1}, {8, -1}, {5, -1}};
    for (int s = 0; s < ORB_NSTAGES; ++s)
        for (int k = 0; k < 2; ++k) {
            if (from[s][k] < 0) continue;
            float ms = 0.f;
            ORB_CUDA(cudaEventElapsedTime(&ms, ev[from[s][k]], ev[to[s][k]]));
            c->prof_ms[s] += ms;
        }
    c->prof_calls++;
    c->prof_total_frames += c->prof_frames[slot];
    c->prof_pending[slot] = false;
    return ORB_OK;
}

extern "C" {

int orb_create(orb_ctx** out, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device,
               int max_batch) {
    if (!out) return ORB_ERR_INVALID;
    *out = nullptr;
    if (nfeatures <= 0 || nlevels <= 0 || nlevels > ORB_MAX_LEVELS || !(scale_factor > 1.0f) || ini_th <= 0 ||
        min_th <= 0 || ini_th > 254 || min_th > ini_th || max_batch <= 0 || max_batch > 65535) {
        orb_set_error("orb_create: invalid parameters");
        return ORB_ERR_INVALID;
    }
    orb_ctx* c = new orb_ctx;
    c->device = device; c->nfeatures = nfeatures; c->nlevels = nlevels; c->ini_th = ini_th; c->min_th = min_th;
    c->scale_factor = scale_factor;  // stored as double like the reference member (ORBextractor.h:99)
    c->max_batch = max_batch;
    // ORBextractor.cc:423-455
    c->scale[0] = 1.f; c->sigma2[0] = 1.f;
    for (int i = 1; i < nlevels; ++i) {
        c->scale[i] = (float)(c->scale[i - 1] * c->scale_factor);
        c->sigma2[i] = c->scale[i] * c->scale[i];
    }
    for (int i = 0; i < nlevels; ++i) { c->inv_scale[i] = 1.f / c->scale[i]; c->inv_sigma2[i] = 1.f / c->sigma2[i]; }
    const float factor = (float)(1.0f / c->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        c->quota[l] = cv_round_f(desired);
        sum += c->quota[l];
        desired *= factor;
    }
    c->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // device side is created lazily so that a context can be constructed (and its tables read) without a GPU
    *out = c;
    return ORB_OK;
}

static int ensure_device(orb_ctx* c) {
    int rc = check_device();
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaSetDevice(c->device));
    if (!c->stream) {
        ORB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    if (!c->d_mom_tab) {
        // IC_Angle (ORBextractor.cc:77-104) as word dot products: for patch alignment a, item = row*9 + word holds the
        // signed u-weights of the word's 4 bytes (0 outside the circular patch, umax from :463-478) and the 0/1 mask
        static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        // (second word: the row offset v = r - 15 times the mask, as signed bytes, so that m01 is a dot product too).  33 rows of
        // 9 words: rows 31 and 32 are zero (the kernel walks 11 steps of 3 rows)
        std::vector<uint2> tab(4 * 297);
        for (int a = 0; a < 4; ++a)
            for (int item = 0; item < 297; ++item) {
                const int r = item / 9, wi = item % 9, v = r - ORB_HALF_PATCH;
                unsigned wu = 0, wv = 0;
                for (int b = 0; b < 4; ++b) {
                    const int u = 4 * wi + b - a - ORB_HALF_PATCH;
                    if (r <= 30 && abs(u) <= umax[abs(v)]) { wu |= (unsigned)(uint8_t)(int8_t)u << (8 * b); wv |= (unsigned)(uint8_t)(int8_t)v << (8 * b); }
                }
                tab[a * 297 + item] = make_uint2(wu, wv);
            }
        // behind it: the rBRIEF pattern (ORBextractor.cc:150-408) as float4[8][32] = (x0, y0, x1, y1) of bit k of descriptor byte i at
        // [k * 32 + i] — the layout orient_describe_kernel keeps in shared memory, fetched with one coalesced 16-byte load per thread
        static const int pattern[1024] = {
#include "orb_pattern_31.inc"
        };
        static_assert(sizeof(uint2) * 4 * 297 % 16 == 0, "the pattern table must start 16-byte aligned");
        std::vector<float> pat(4 * 256);
        for (int k = 0; k < 8; ++k)
            for (int i = 0; i < 32; ++i)
                for (int e = 0; e < 4; ++e) pat[4 * (k * 32 + i) + e] = (float)pattern[i * 32 + 4 * k + e];
        ORB_CUDA(cudaMalloc(&c->d_mom_tab, sizeof(uint2) * tab.size() + sizeof(float) * pat.size()));
        ORB_CUDA(cudaMemcpy(c->d_mom_tab, tab.data(), sizeof(uint2) * tab.size(), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_mom_tab + tab.size(), pat.data(), sizeof(float) * pat.size(), cudaMemcpyHostToDevice));
    }
    if (!c->st_h2d) {
        ORB_CUDA(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
        ORB_CUDA(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
        ORB_CUDA(cudaStreamCreateWithFlags(&c->st_c2, cudaStreamNonBlocking));
        // measured on B200 (512 x 640x480): running border -> blur on a second stream next to FAST -> quadtree changes
        // nothing (154.4 k vs 154.1 k frames/s: the FAST grid alone fills the chip), so the serial order is the default
        c->overlap = getenv("ORB_B200_OVERLAP") != nullptr;
        for (int i = 0; i < 2; ++i) {
            ORB_CUDA(cudaStreamCreateWithFlags(&c->st_aux[i], cudaStreamNonBlocking));
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_pyr[i], cudaEventDisableTiming));
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_blur[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < ORB_PIPE_SLOTS; ++i) {
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            ORB_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
        }
    }
    return ORB_OK;
}

void orb_destroy(orb_ctx* c) {
    if (!c) return;
    if (orb_device_count() > 0) {
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        free_geometry_buffers(c);
        cudaFree(c->d_mom_tab); cudaFree(c->d_scratch); cudaFreeHost(c->h_scratch);
        cudaFree(c->d_fast_stats); cudaFreeHost(c->h_fast_stats);
        if (c->ev_fast_stats) cudaEventDestroy(c->ev_fast_stats);
        if (c->prof_ev[0][0])
            for (int r = 0; r < ORB_PROF_RING; ++r)
                for (int s = 0; s < ORB_PROF_EVENTS; ++s) cudaEventDestroy(c->prof_ev[r][s]);
        if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
        if (c->st_h2d) {
            cudaStreamDestroy(c->st_h2d); cudaStreamDestroy(c->st_d2h); cudaStreamDestroy(c->st_c2);
            for (int i = 0; i < 2; ++i) { cudaStreamDestroy(c->st_aux[i]); cudaEventDestroy(c->ev_pyr[i]); cudaEventDestroy(c->ev_blur[i]); }
            if (c->ev_head) { cudaEventDestroy(c->ev_head); cudaEventDestroy(c->ev_early); }
            for (int i = 0; i < ORB_PIPE_SLOTS; ++i) { cudaEventDestroy(c->ev_in[i]); cudaEventDestroy(c->ev_done[i]); cudaEventDestroy(c->ev_out[i]); }
        }
    }
    delete c;
}

int orb_get_tables(orb_ctx* c, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* per_level) {
    if (!c) return ORB_ERR_INVALID;
    for (int i = 0; i < c->nlevels; ++i) {
        if (scale) scale[i] = c->scale[i];
        if (inv_scale) inv_scale[i] = c->inv_scale[i];
        if (sigma2) sigma2[i] = c->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = c->inv_sigma2[i];
        if (per_level) per_level[i] = c->quota[i];
    }
    return ORB_OK;
}

int orb_max_keypoints(orb_ctx* c) {
    if (!c) return ORB_ERR_INVALID;
    // list length after DistributeOctTree <= max(N_l + 2, 4 * nIni); nIni <= 16 for any sane aspect ratio
    int tot = 0;
    for (int l = 0; l < c->nlevels; ++l) tot += std::max(c->quota[l] + 3, 64);
    return tot;
}

int orb_set_stream(orb_ctx* c, void* s) {
    if (!c) return ORB_ERR_INVALID;
    int rc = ensure_device(c);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)s;
    if (!c->stream) {
        ORB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return ORB_OK;
}

int orb_sync(orb_ctx* c) {
    if (!c) return ORB_ERR_INVALID;
    if (!c->stream) return ORB_OK;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

int64_t orb_launch_count(orb_ctx* c) { return c ? c->launches : 0; }

int orb_profile_enable(orb_ctx* c, int enable) {
    if (!c) return ORB_ERR_INVALID;
    int rc = ensure_device(c);
    if (rc != ORB_OK) return rc;
    if (enable && !c->prof_ev[0][0]) {
        for (int r = 0; r < ORB_PROF_RING; ++r)
            for (int s = 0; s < ORB_PROF_EVENTS; ++s) ORB_CUDA(cudaEventCreate(&c->prof_ev[r][s]));
    }
    c->profile = enable != 0;
    return ORB_OK;
}

int orb_profile_read(orb_ctx* c, double* stage_ms, int64_t* calls, int64_t* frames, int reset) {
    if (!c) return ORB_ERR_INVALID;
    if (c->prof_ev[0][0]) {
        ORB_CUDA(cudaSetDevice(c->device));
        for (int r = 0; r < ORB_PROF_RING; ++r)
            if (c->prof_pending[r]) { int rc = orb_profile_harvest(c, r); if (rc != ORB_OK) return rc; }
    }
    if (stage_ms) for (int s = 0; s < ORB_NSTAGES; ++s) stage_ms[s] = c->prof_ms[s];
    if (calls) *calls = c->prof_calls;
    if (frames) *frames = c->prof_total_frames;
    if (reset) { for (int s = 0; s < ORB_NSTAGES; ++s) c->prof_ms[s] = 0; c->prof_calls = 0; c->prof_total_frames = 0; }
    return ORB_OK;
}


// One shared-memory carve-out for every kernel of the chain while chunks of the host pipeline run on several streams.
// Measured (profiles/r1_l_carveout.txt): kernels with different carve-out needs (FAST: 190 KB of shared memory per SM,
// the others: none, all of it L1) cannot share an SM — it has to drain before its L1 / shared-memory split changes — so
// concurrent chunks, or two extractor instances on two threads, slowed each other down (2 x 256 frames concurrently:
// 159 k frames/s against 171 k back to back).  With the same 58 % split everywhere two concurrent callers reach 164 k
// frames/s instead of 136 k; the device-resident path keeps the per-kernel default (its launches never overlap, and
// orient_describe lives on a large L1: 0.53 ms -> 1.33 ms with a carve-out of 72 % or more).
// pct: 0..100, or -1 = the driver's per-kernel default.  ORB_B200_CARVEOUT overrides both uses.
static void chain_carveout(orb_ctx* c, int pct) {
    static const int forced = [] { const char* e = getenv("ORB_B200_CARVEOUT"); return e ? atoi(e) : -2; }();
    if (forced >= -1 && forced <= 100) pct = forced;
    static std::atomic<int> current[64];   // per device: last value set + 3 (0 = nothing set yet)
    if (current[c->device & 63].exchange(pct + 3) == pct + 3) return;
    orb_carveout_pyramid(pct); orb_carveout_blur(pct); orb_carveout_fast(pct); orb_carveout_extract(pct);
}

int orb_extract_batch_device(orb_ctx* c, const uint8_t* d_imgs, int nframes, int w, int h, size_t row_stride,
                             size_t frame_stride, orb_kp* d_kps, uint8_t* d_desc, int cap, int32_t* d_n_out) {
    return orb_extract_batch_device_pix(c, d_imgs, ORB_PIX_GRAY8, nframes, w, h, row_stride, frame_stride, d_kps, d_desc, cap, d_n_out);
}

int orb_extract_batch_device_pix(orb_ctx* c, const uint8_t* d_imgs, int fmt, int nframes, int w, int h, size_t row_stride,
                                 size_t frame_stride, orb_kp* d_kps, uint8_t* d_desc, int cap, int32_t* d_n_out) {
    if (!c || !d_imgs || !d_kps || !d_desc || !d_n_out || nframes < 0 || cap <= 0) return ORB_ERR_INVALID;
    if (fmt < ORB_PIX_GRAY8 || fmt > ORB_PIX_RGBA8) return ORB_ERR_INVALID;
    if (nframes == 0 || w <= 0 || h <= 0) return ORB_OK;  // empty image => silent return (ORBextractor.cc:1086)
    if (nframes > c->max_batch) { orb_set_error("nframes %d > max_batch %d", nframes, c->max_batch); return ORB_ERR_CAPACITY; }
    if (row_stride < (size_t)w * orb_pix_channels(fmt)) return ORB_ERR_INVALID;
    int rc = ensure_device(c);
    if (rc != ORB_OK) return rc;
    rc = build_geometry(c, w, h);
    if (rc != ORB_OK) return rc;
    c->last_frames = nframes;
    chain_carveout(c, -1);
    // (measured: cutting a resident batch into sub-batches on two streams is SLOWER, 147k -> 130k frames/s at 512
    // frames; the big launches already fill the chip.  Only the host pipeline below alternates streams.)
    return orb_launch_extract(c, d_imgs, fmt, nframes, 0, row_stride, frame_stride, d_kps, d_desc, cap, d_n_out, c->stream);
}

// Host buffers in, host buffers out.  The batch is cut into chunks of ORB_PIPE_CHUNK frames that flow through three
// streams (H2D copy -> kernels -> D2H copy), so the PCIe transfers of neighbouring chunks overlap the kernels.
// Pinned caller buffers are used directly; pageable ones go through the context's pinned mirrors.
int orb_extract_batch(orb_ctx* c, const uint8_t* imgs, int nframes, int w, int h, size_t row_stride, size_t frame_stride,
                      orb_kp* kps, uint8_t* desc, int cap, int32_t* n_out) {
    return orb_extract_batch_pix(c, imgs, ORB_PIX_GRAY8, nframes, w, h, row_stride, frame_stride, kps, desc, cap, n_out);
}

// One frame, pageable (or any) host buffers: copy the image into the pinned staging buffer, launch the captured graph,
// wait, copy the results out.  Returns 1 when the graph path is not applicable (the caller takes the pipeline path).
static int extract_one_graph(orb_ctx* c, const uint8_t* img, int fmt, int w, int h, size_t row_stride, orb_kp* kps, uint8_t* desc,
                             int cap, int32_t* n_out) {
    if (c->use_graph < 0) { const char* e = getenv("ORB_B200_GRAPH"); c->use_graph = (e && atoi(e) == 0) ? 0 : 1; }
    if (!c->use_graph || c->profile) return 1;
    // the first call of a geometry / format runs without a graph: it sets the kernels' attributes and sizes the buffers
    if (c->graph_warm_w != w || c->graph_warm_h != h || c->graph_warm_fmt != fmt) {
        c->graph_warm_w = w; c->graph_warm_h = h; c->graph_warm_fmt = fmt;
        return 1;
    }
    const int ch = orb_pix_channels(fmt);
    const size_t rbytes = (size_t)w * ch, fbytes = rbytes * h;
    const int dcap = std::min(cap, c->g.total_kp_slots);
    int rc = ensure_outputs(c, std::max(dcap, c->out_cap), true);
    if (rc != ORB_OK) return rc;
    rc = ensure_input_staging(c, true, ch);
    if (rc != ORB_OK) return rc;
    const int ocap = c->out_cap, rows = std::min(ocap, cap);
    const orb_ctx::GraphKey key = {fmt, w, h, rows, ocap, c->d_in, c->h_in, c->h_kps, c->h_desc, c->d_kps_out};
    if (!c->graph_exec || memcmp(&key, &c->graph_key, sizeof(key)) != 0) {
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        ORB_CUDA(cudaStreamSynchronize(c->stream));
        const long long l0 = c->launches;
        const bool ov = c->overlap;
        cudaGraph_t graph = nullptr;
        ORB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = cudaMemcpyAsync(c->d_in, c->h_in, fbytes, cudaMemcpyHostToDevice, c->stream);
        c->overlap = true;   // inside the graph the two independent chains are parallel branches
        if (e == cudaSuccess)
            rc = orb_launch_extract(c, c->d_in, fmt, 1, 0, rbytes, fbytes, c->d_kps_out, c->d_desc_out, ocap, c->d_n_out, c->stream);
        c->overlap = ov;
        if (e == cudaSuccess && rc == ORB_OK) e = cudaMemcpyAsync(c->h_n, c->d_n_out, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && rc == ORB_OK) e = cudaMemcpyAsync(c->h_kps, c->d_kps_out, sizeof(orb_kp) * (size_t)rows, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && rc == ORB_OK) e = cudaMemcpyAsync(c->h_desc, c->d_desc_out, (size_t)32 * rows, cudaMemcpyDeviceToHost, c->stream);
        const cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
        if (e != cudaSuccess || e2 != cudaSuccess || rc != ORB_OK || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            c->use_graph = 0;   // capture not possible here: keep using the pipeline path
            c->launches = l0;
            return 1;
        }
        e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); c->graph_exec = nullptr; c->use_graph = 0; c->launches = l0; return 1; }
        c->graph_launches = c->launches - l0;
        c->launches = l0;
        c->graph_key = key;
    }
    if (row_stride == rbytes) memcpy(c->h_in, img, fbytes);
    else for (int y = 0; y < h; ++y) memcpy(c->h_in + (size_t)y * rbytes, img + (size_t)y * row_stride, rbytes);
    ORB_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
    c->launches += c->graph_launches;
    c->last_frames = 1;
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    const int n = c->h_n[0];
    n_out[0] = n;
    const int m = std::min(n, rows);
    memcpy(kps, c->h_kps, sizeof(orb_kp) * (size_t)m);
    memcpy(desc, c->h_desc, (size_t)32 * m);
    if (n > cap) { orb_set_error("frame 0 produced %d keypoints, capacity %d", n, cap); return ORB_ERR_CAPACITY; }
    return ORB_OK;
}

int orb_extract_batch_pix(orb_ctx* c, const uint8_t* imgs, int fmt, int nframes, int w, int h, size_t row_stride,
                          size_t frame_stride, orb_kp* kps, uint8_t* desc, int cap, int32_t* n_out) {
    if (!c || !n_out || nframes < 0 || cap < 0) return ORB_ERR_INVALID;
    if (fmt < ORB_PIX_GRAY8 || fmt > ORB_PIX_RGBA8) return ORB_ERR_INVALID;
    const int ch = orb_pix_channels(fmt);
    for (int f = 0; f < nframes; ++f) n_out[f] = 0;
    if (nframes == 0 || !imgs || w <= 0 || h <= 0) return ORB_OK;  // ORBextractor.cc:1086-1087
    if (!kps || !desc || cap == 0 || row_stride < (size_t)w * ch) return ORB_ERR_INVALID;
    int rc = ensure_device(c);
    if (rc != ORB_OK) return rc;
    rc = build_geometry(c, w, h);
    if (rc != ORB_OK) return rc;
    if (nframes == 1) {   // latency path
        rc = extract_one_graph(c, imgs, fmt, w, h, row_stride, kps, desc, cap, n_out);
        if (rc != 1) return rc;
    }
    const size_t rbytes = (size_t)w * ch;                 // bytes of one tight row
    const bool tight = row_stride == rbytes && frame_stride == rbytes * h;
    const bool in_direct = tight && is_pinned(imgs);
    const bool out_direct = is_pinned(kps) && is_pinned(desc);
    const int dcap = std::min(cap, c->g.total_kp_slots);
    rc = ensure_outputs(c, std::max(dcap, c->out_cap), !out_direct);
    if (rc != ORB_OK) return rc;
    rc = ensure_input_staging(c, !in_direct, ch);
    if (rc != ORB_OK) return rc;
    const int ocap = c->out_cap;                        // device row length (>= dcap)
    chain_carveout(c, 58);
    static const int chunk_frames = [] {               // frames per pipeline chunk (tunable for experiments)
        const char* e = getenv("ORB_B200_PIPE_CHUNK");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : ORB_PIPE_CHUNK;
    }();
    const size_t fbytes = rbytes * h;
    int status = ORB_OK;
    // ORB_B200_PIPE_TRACE=1: per-chunk timeline (H2D start / end, kernels end, D2H end, ms since the call's start) on stderr
    static const bool trace = getenv("ORB_B200_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t0ev = nullptr;
    if (trace) { cudaEventCreate(&t0ev); cudaEventRecord(t0ev, c->st_h2d); }
    // the caller's stream (c->stream) may have pending work that produced or still reads our buffers
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    for (int b0 = 0; b0 < nframes; b0 += c->max_batch) {
        const int B = std::min(c->max_batch, nframes - b0);
        // chunk schedule: two half-size chunks first (the kernels start after 1/16 of a 512-frame batch has arrived), then
        // full chunks, and the last full chunk's worth of frames as 1/2 + 1/4 + 1/4: the H2D stream is the bottleneck, so
        // the call ends when the kernels + D2H of the LAST chunk are through — the smaller it is, the shorter that tail
        static const bool taper = [] { const char* e = getenv("ORB_B200_PIPE_TAPER"); return !(e && atoi(e) == 0); }();   // measured with the common carve-out: +2 % for one caller
        std::vector<int> cf0, cF;
        {
            int f = 0, k = 0;
            while (f < B) {
                int want = (k < 2 && B > chunk_frames) ? std::max(chunk_frames / 2, 1) : chunk_frames;
                const int left = B - f;
                if (taper && B >= 4 * chunk_frames && chunk_frames >= 32) {
                    if (left <= chunk_frames / 2) want = std::max(chunk_frames / 4, 1);
                    else if (left <= chunk_frames) want = chunk_frames / 2;
                }
                const int F = std::min(want, left);
                cf0.push_back(f); cF.push_back(F);
                f += F; ++k;
            }
        }
        const int nchunks = (int)cf0.size();
        c->last_frames = B;
        auto finish_chunk = [&](int k) -> int {       // host side of chunk k once its D2H has landed
            const int f0 = cf0[k], F = cF[k];
            ORB_CUDA(cudaEventSynchronize(c->ev_out[k % ORB_PIPE_SLOTS]));
            for (int f = 0; f < F; ++f) {
                const int n = c->h_n[f0 + f];
                n_out[b0 + f0 + f] = n;
                if (n > cap) { orb_set_error("frame %d produced %d keypoints, capacity %d", b0 + f0 + f, n, cap); status = ORB_ERR_CAPACITY; }
                if (!out_direct) {
                    const int m = std::min(n, std::min(cap, ocap));
                    memcpy(kps + (size_t)(b0 + f0 + f) * cap, c->h_kps + (size_t)(f0 + f) * ocap, sizeof(orb_kp) * m);
                    memcpy(desc + (size_t)(b0 + f0 + f) * cap * 32, c->h_desc + (size_t)(f0 + f) * ocap * 32, (size_t)32 * m);
                }
            }
            return ORB_OK;
        };
        for (int k = 0; k < nchunks; ++k) {
            const int f0 = cf0[k], F = cF[k];
            const int slot = k % ORB_PIPE_SLOTS;
            if (k >= ORB_PIPE_SLOTS) { rc = finish_chunk(k - ORB_PIPE_SLOTS); if (rc != ORB_OK) return rc; }   // frees the slot's events
            // ---- H2D ----
            const uint8_t* src = imgs + (size_t)(b0 + f0) * frame_stride;
            if (!in_direct) {
                uint8_t* stage = c->h_in + (size_t)f0 * fbytes;
                for (int f = 0; f < F; ++f) {
                    const uint8_t* s1 = src + (size_t)f * frame_stride;
                    uint8_t* d1 = stage + (size_t)f * fbytes;
                    if (row_stride == rbytes) memcpy(d1, s1, fbytes);
                    else for (int y = 0; y < h; ++y) memcpy(d1 + (size_t)y * rbytes, s1 + (size_t)y * row_stride, rbytes);
                }
                src = stage;
            }
            uint8_t* d_src = c->d_in + (size_t)f0 * fbytes;
            if (trace) { tev.resize(tev.size() + 4); for (int i = 0; i < 4; ++i) cudaEventCreate(&tev[tev.size() - 4 + i]); cudaEventRecord(tev[tev.size() - 4], c->st_h2d); }
            ORB_CUDA(cudaMemcpyAsync(d_src, src, (size_t)F * fbytes, cudaMemcpyHostToDevice, c->st_h2d));
            ORB_CUDA(cudaEventRecord(c->ev_in[slot], c->st_h2d));
            if (trace) cudaEventRecord(tev[tev.size() - 3], c->st_h2d);
            // ---- kernels: chunks alternate between two compute streams, so the latency-bound small launches of one
            // chunk (upper pyramid levels, quadtree) overlap the issue-bound ones of its neighbour ----
            static const int nstreams = [] { const char* e = getenv("ORB_B200_PIPE_STREAMS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 4) ? v : 4; }();
            cudaStream_t css[4] = {c->stream, c->st_c2, c->st_aux[0], c->st_aux[1]};
            cudaStream_t cs = css[(c->overlap ? (k & 1) : k % nstreams)];
            ORB_CUDA(cudaStreamWaitEvent(cs, c->ev_in[slot], 0));
            rc = orb_launch_extract(c, d_src, fmt, F, f0, rbytes, fbytes, c->d_kps_out + (size_t)f0 * ocap,
                                    c->d_desc_out + (size_t)f0 * ocap * 32, ocap, c->d_n_out + f0, cs);
            if (rc != ORB_OK) return rc;
            ORB_CUDA(cudaEventRecord(c->ev_done[slot], cs));
            if (trace) cudaEventRecord(tev[tev.size() - 2], cs);
            // ---- D2H ----
            ORB_CUDA(cudaStreamWaitEvent(c->st_d2h, c->ev_done[slot], 0));
            ORB_CUDA(cudaMemcpyAsync(c->h_n + f0, c->d_n_out + f0, sizeof(int) * F, cudaMemcpyDeviceToHost, c->st_d2h));
            const int rows = std::min(ocap, cap);       // keypoint slots copied per frame
            if (out_direct) {
                ORB_CUDA(cudaMemcpy2DAsync(kps + (size_t)(b0 + f0) * cap, sizeof(orb_kp) * (size_t)cap, c->d_kps_out + (size_t)f0 * ocap,
                                           sizeof(orb_kp) * (size_t)ocap, sizeof(orb_kp) * (size_t)rows, F, cudaMemcpyDeviceToHost, c->st_d2h));
                ORB_CUDA(cudaMemcpy2DAsync(desc + (size_t)(b0 + f0) * cap * 32, (size_t)32 * cap, c->d_desc_out + (size_t)f0 * ocap * 32,
                                           (size_t)32 * ocap, (size_t)32 * rows, F, cudaMemcpyDeviceToHost, c->st_d2h));
            } else {
                ORB_CUDA(cudaMemcpyAsync(c->h_kps + (size_t)f0 * ocap, c->d_kps_out + (size_t)f0 * ocap, sizeof(orb_kp) * (size_t)F * ocap,
                                         cudaMemcpyDeviceToHost, c->st_d2h));
                ORB_CUDA(cudaMemcpyAsync(c->h_desc + (size_t)f0 * ocap * 32, c->d_desc_out + (size_t)f0 * ocap * 32, (size_t)32 * F * ocap,
                                         cudaMemcpyDeviceToHost, c->st_d2h));
            }
            ORB_CUDA(cudaEventRecord(c->ev_out[slot], c->st_d2h));
            if (trace) cudaEventRecord(tev[tev.size() - 1], c->st_d2h);
        }
        for (int k = std::max(0, nchunks - ORB_PIPE_SLOTS); k < nchunks; ++k) { rc = finish_chunk(k); if (rc != ORB_OK) return rc; }
        ORB_CUDA(cudaStreamSynchronize(c->stream));
        ORB_CUDA(cudaStreamSynchronize(c->st_c2));
        ORB_CUDA(cudaStreamSynchronize(c->st_aux[0]));
        ORB_CUDA(cudaStreamSynchronize(c->st_aux[1]));
    }
    if (trace) {
        for (size_t k = 0; k < tev.size() / 4; ++k) {
            float t[4];
            for (int i = 0; i < 4; ++i) { cudaEventElapsedTime(&t[i], t0ev, tev[4 * k + i]); cudaEventDestroy(tev[4 * k + i]); }
            fprintf(stderr, "[pipe] chunk %2zu  h2d %.3f-%.3f  kernels done %.3f  d2h done %.3f ms\n", k, t[0], t[1], t[2], t[3]);
        }
        cudaEventDestroy(t0ev);
    }
    return status;
}

int orb_extract(orb_ctx* c, const uint8_t* img, int w, int h, size_t stride, orb_kp* kps, uint8_t* desc, int cap,
                int* n_out) {
    if (!n_out) return ORB_ERR_INVALID;
    int32_t n = 0;
    const int rc = orb_extract_batch(c, img, (img && w > 0 && h > 0) ? 1 : 0, w, h, stride, (size_t)stride * h, kps, desc,
                                     cap, &n);
    *n_out = n;
    return rc;
}

int orb_level_dims(orb_ctx* c, int level, int* w, int* h) {
    if (!c || !c->have_geom || level < 0 || level >= c->nlevels) return ORB_ERR_INVALID;
    if (w) *w = c->g.lv[level].w;
    if (h) *h = c->g.lv[level].h;
    return ORB_OK;
}

int orb_pyramid_level(orb_ctx* c, int frame, int level, uint8_t* dst, size_t dst_stride) {
    if (!c || !c->have_geom || !dst || level < 0 || level >= c->nlevels || frame < 0 || frame >= c->last_frames)
        return ORB_ERR_INVALID;
    const LevelGeom& L = c->g.lv[level];
    const int bw = L.w + 2 * ORB_EDGE;
    if (dst_stride < (size_t)bw) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaMemcpy2DAsync(dst, dst_stride, c->d_pyr + L.base + (long long)frame * L.frame_stride + (ORB_XOFF - ORB_EDGE), L.pitch, bw, L.rows,
                               cudaMemcpyDeviceToHost, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

/* every level of one frame with ONE synchronisation: the bordered buffer of level l, (h_l + 38) rows of (w_l + 38) bytes (the
 * reference's own layout: step = w + 38), lands at dst + offset[l]; pitch[l] = w_l + 38 */
int orb_pyramid_levels(orb_ctx* c, int frame, uint8_t* dst, size_t dst_bytes, size_t* offset, size_t* pitch, size_t* needed) {
    if (!c || !c->have_geom || frame < 0 || frame >= c->last_frames) return ORB_ERR_INVALID;
    size_t total = 0;
    for (int l = 0; l < c->nlevels; ++l) total += (size_t)(c->g.lv[l].w + 2 * ORB_EDGE) * c->g.lv[l].rows;
    if (needed) *needed = total;
    if (!dst) return needed ? ORB_OK : ORB_ERR_INVALID;
    if (dst_bytes < total || !offset || !pitch) return ORB_ERR_CAPACITY;
    ORB_CUDA(cudaSetDevice(c->device));
    size_t off = 0;
    for (int l = 0; l < c->nlevels; ++l) {
        const LevelGeom& L = c->g.lv[l];
        const size_t bw = (size_t)L.w + 2 * ORB_EDGE;
        ORB_CUDA(cudaMemcpy2DAsync(dst + off, bw, c->d_pyr + L.base + (long long)frame * L.frame_stride + (ORB_XOFF - ORB_EDGE), L.pitch, bw, L.rows,
                                   cudaMemcpyDeviceToHost, c->stream));
        offset[l] = off;
        pitch[l] = bw;
        off += bw * L.rows;
    }
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

int orb_debug_blurred(orb_ctx* c, int frame, int level, uint8_t* dst, size_t dst_stride) {
    if (!c || !c->have_geom || !dst || level < 0 || level >= c->nlevels || frame < 0 || frame >= c->last_frames)
        return ORB_ERR_INVALID;
    const LevelGeom& L = c->g.lv[level];
    if (dst_stride < (size_t)L.w) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaMemcpy2DAsync(dst, dst_stride, c->d_blur + L.bbase + (long long)frame * L.bframe_stride, L.bpitch, L.w, L.h,
                               cudaMemcpyDeviceToHost, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

int orb_debug_raw_corners(orb_ctx* c, int frame, int level, float* xyr, int cap, int* n_out) {
    if (!c || !c->have_geom || !n_out || level < 0 || level >= c->nlevels || frame < 0 || frame >= c->last_frames)
        return ORB_ERR_INVALID;
    const LevelGeom& L = c->g.lv[level];
    ORB_CUDA(cudaSetDevice(c->device));
    int n = 0;
    ORB_CUDA(cudaMemcpyAsync(&n, c->d_corner_count + frame * c->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    n = std::min(n, L.corner_cap);
    *n_out = n;
    if (!xyr || n == 0) return ORB_OK;
    std::vector<unsigned long long> rec(n);
    ORB_CUDA(cudaMemcpyAsync(rec.data(), c->d_corners + L.corner_base + (long long)frame * L.corner_cap,
                             sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    std::sort(rec.begin(), rec.end(), [](unsigned long long a, unsigned long long b) { return corner_order(a) < corner_order(b); });
    for (int i = 0; i < std::min(n, cap); ++i) {
        xyr[3 * i] = (float)corner_x(rec[i]); xyr[3 * i + 1] = (float)corner_y(rec[i]); xyr[3 * i + 2] = (float)corner_score(rec[i]);
    }
    return n > cap ? ORB_ERR_CAPACITY : ORB_OK;
}

int orb_debug_tie_counts(orb_ctx* c, int frame, int32_t* ties /*[nlevels]*/) {
    if (!c || !c->have_geom || !ties || frame < 0 || frame >= c->last_frames) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaMemcpyAsync(ties, c->d_corner_count + (size_t)c->max_batch * c->nlevels + frame * c->nlevels,
                             sizeof(int) * c->nlevels, cudaMemcpyDeviceToHost, c->stream));
    ORB_CUDA(cudaStreamSynchronize(c->stream));
    return ORB_OK;
}

}  // extern "C"
