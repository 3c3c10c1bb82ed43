/* synth_int.c — the integer-only, counter-based synthetic frame generator of SURVEY.md §8d in C: byte-identical to
 * orb_slam_2_ros_b200.synth.synth_frame_int (tests/test_synth_int.py builds this file with gcc and compares).  No RNG state,
 * no libm, no OpenCV: every random number is splitmix64(seed * 2^32 + counter).
 *   void synth_frame_int(uint32_t seed, int w, int h, int n_rect, int noise, uint8_t* out)   // out: h rows of w bytes */
#include <stdint.h>
#include <stdlib.h>

static uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct { uint64_t base, ctr; } Rnd;
static int64_t rnd(Rnd* r, int64_t n) { r->ctr++; return (int64_t)((sm64(r->base + r->ctr) >> 11) % (uint64_t)n); }

static void fill(int64_t* img, int w, int h, int x0, int y0, int rw, int rh, int64_t v) {
    for (int y = y0; y < y0 + rh && y < h; ++y)
        for (int x = x0; x < x0 + rw && x < w; ++x) img[(size_t)y * w + x] = v;
}

void synth_frame_int(uint32_t seed, int w, int h, int n_rect, int noise, uint8_t* out) {
    Rnd r = {(uint64_t)seed << 32, 0};
    enum { GW = 6, GH = 5 };
    int64_t coarse[GH + 1][GW + 1];
    for (int y = 0; y <= GH; ++y) for (int x = 0; x <= GW; ++x) coarse[y][x] = 40 + rnd(&r, 176);
    int64_t* img = (int64_t*)malloc(sizeof(int64_t) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const int64_t ys = ((int64_t)y * GH * 256) / h, yi = ys >> 8, yf = ys & 255;
        for (int x = 0; x < w; ++x) {
            const int64_t xs = ((int64_t)x * GW * 256) / w, xi = xs >> 8, xf = xs & 255;
            const int64_t top = coarse[yi][xi] * (256 - xf) + coarse[yi][xi + 1] * xf;
            const int64_t bot = coarse[yi + 1][xi] * (256 - xf) + coarse[yi + 1][xi + 1] * xf;
            img[(size_t)y * w + x] = (top * (256 - yf) + bot * yf) >> 16;
        }
    }
    for (int k = 0; k < n_rect; ++k) {
        const int x0 = (int)rnd(&r, w), y0 = (int)rnd(&r, h), rw = 6 + (int)rnd(&r, 64), rh = 6 + (int)rnd(&r, 64);
        const int64_t v = rnd(&r, 256);
        fill(img, w, h, x0, y0, rw, rh, v);
    }
    const int tw = w / 3, th = h / 2;
    for (int by = 0; by < th; by += 4)
        for (int bx = 0; bx < tw; bx += 4) fill(img, w, h, bx, by, 4, 4, rnd(&r, 256));
    fill(img, w, h, w - w / 4, h - h / 4, w / 4 + 1, h / 4 + 1, 128);
    const int lx = w / 4, ly = h / 4;
    for (int by = h - ly; by < h; by += 12)
        for (int bx = 0; bx < lx; bx += 12) fill(img, w, h, bx, by, 12, 12, 100 + ((bx / 12 + by / 12) & 1) * (8 + rnd(&r, 13)));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int64_t v = img[(size_t)y * w + x];
            if (noise > 0 && !(y >= h - h / 4 && x >= w - w / 4)) {
                const uint64_t idx = (uint64_t)y * w + x + (((uint64_t)seed << 32) + (1ull << 31));
                v += (int64_t)((sm64(idx) >> 11) % (uint64_t)(2 * noise + 1)) - noise;
            }
            out[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(img);
}
