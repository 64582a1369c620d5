// car_spans.cuh -- pygame 1.9 draw_fillpoly scanline rule and the per-track span tables of the road polygons.
// Shared by the track generator (car_physics.cu: tables built once per track) and the rasteriser (car_raster.cu).
#pragma once
#include "car_common.cuh"

namespace crl {

// C integer division a / b (b > 0, truncation towards zero).  For |a| < 2^24 the correctly rounded float quotient of
// two integers truncates to the exact answer (a non-integer quotient is at least 1/b away from an integer, the
// rounding error is below |a/b| * 2^-24), which is several times cheaper than the emulated 32-bit division.
__device__ __forceinline__ int cdiv_trunc(int a, int b) {
    if (abs(a) < (1 << 24)) return (int)__fdiv_rn((float)a, (float)b);
    return a / b;
}

// pygame 1.9 draw_fillpoly, one scanline: x spans (inclusive) of polygon (vx, vy)[n] at row V.
// Up to two spans (outlines with <= 8 vertices used here never give more); empty span = (1, 0).
__device__ __forceinline__ short4 scanline_spans(const short* vx, const short* vy, int n, int V, int maxy) {
    int xs[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    int m = 0;
    int yp = vy[n - 1], xp = vx[n - 1];                   // previous vertex (edge i runs from vertex i - 1 to vertex i)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i >= n) break;
        const int yc = vy[i], xc = vx[i];
        int y1 = yp, y2 = yc, x1 = xp, x2 = xc;
        yp = yc; xp = xc;
        if (y1 > y2) { int t = y1; y1 = y2; y2 = t; t = x1; x1 = x2; x2 = t; }
        if (y1 != y2 && ((V >= y1 && V < y2) || (V == maxy && V > y1 && V <= y2))) {
            const int x = cdiv_trunc((V - y1) * (x2 - x1), y2 - y1) + x1;      // C integer division
            if (m == 0) xs[0] = x; else if (m == 1) xs[1] = x; else if (m == 2) xs[2] = x; else if (m == 3) xs[3] = x;
            ++m;
        }
    }
    // sort (unused slots hold INT_MAX): 4-element network
#define CSWAP(a, b) { const int lo_ = min(xs[a], xs[b]), hi_ = max(xs[a], xs[b]); xs[a] = lo_; xs[b] = hi_; }
    CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
#undef CSWAP
    m = min(m, 4);
    short4 r = make_short4(1, 0, 1, 0);
    if (m >= 2) { r.x = (short)xs[0]; r.y = (short)xs[1]; }
    if (m >= 4) { r.z = (short)xs[2]; r.w = (short)xs[3]; }
    return r;
}

// Scanline span tables of every road polygon (tile, kerb) of the track in `slot`, in road-map pixels: they depend on the
// track only, so the generator scans them once per track rather than the render kernel once per frame.  `nthreads`
// threads, thread `tid` of them; one (tile, table row) per thread and pass.
__device__ inline void build_tile_spans(const CarDev& p, int slot, int n_track, int tid, int nthreads) {
    const CarTile* tiles = p.tiles + (size_t)slot * CAR_MAX_TRACK;
    short4* out = p.tile_spans + (size_t)slot * CAR_MAX_TRACK * CAR_SPAN_ROWS;
    for (int item = tid; item < n_track * CAR_SPAN_ROWS; item += nthreads) {
        const int t = item / CAR_SPAN_ROWS, slot_r = item % CAR_SPAN_ROWS;
        const bool kerb = slot_r >= CAR_SPAN_TILE_ROWS;
        const int r = kerb ? slot_r - CAR_SPAN_TILE_ROWS : slot_r;
        const CarTile* T = tiles + t;
        if (kerb && !(T->flags & 2)) continue;
        const short* vx = kerb ? T->kmx : T->mx;
        const short* vy = kerb ? T->kmy : T->my;
        const int n = kerb ? 4 : 5;
        int miny = vy[0], maxy = vy[0];
        for (int i = 1; i < n; ++i) { miny = min(miny, (int)vy[i]); maxy = max(maxy, (int)vy[i]); }
        if (r > maxy - miny) continue;
        out[(size_t)t * CAR_SPAN_ROWS + slot_r] = scanline_spans(vx, vy, n, miny + r, maxy);
    }
}

}  // namespace crl
