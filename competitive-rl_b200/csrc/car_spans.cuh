// car_spans.cuh -- pygame 1.9 draw_fillpoly scanline rule and the per-track road map painted with it.
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
    int yp = vy[0], xp = vx[0];                           // previous vertex (edge i runs from vertex i - 1 to vertex i): the last one
#pragma unroll
    for (int i = 1; i < 8; ++i)                           // (static indices: a register array stays in registers)
        if (i == n - 1) { yp = vy[i]; xp = vx[i]; }
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

// ------------------------------------------------------------------------------------------------
// The road map.  The reference paints every road tile and kerb once per reset into a 10 000 x 10 000 px surface
// (render_road_for_observation_map, car_racing_multi_players.py:732-755) and crops it per frame.  The painted part of that
// surface depends on the track only, so it is kept per track slot as a sparse raster: the 2048 x 2048 px window
// [CAR_MAP_ORIGIN, CAR_MAP_ORIGIN + 2048)^2 around the map centre (+-580 track units; a track stays within +-240) is cut
// into 16 x 16 px blocks, `index[by][bx]` = 0 (nothing painted there), 0xFFFF (block dropped: pool full, flagged) or
// 1 + the block's position in the slot's pool of 256-byte blocks.  A byte is the final gray value of that map pixel: a block
// starts out as the background (grass, or the lighter checker squares of :733-746, which are axis-aligned in the map and the
// same for every track: CarDev::chk) and takes the gray of the last polygon painted over each pixel.
// One warp paints a track: polygons strictly in the reference's paint order (tile n-1 .. 0, each followed by its kerb),
// lanes = scanlines of the polygon (pygame 1.9 draw_fillpoly, above), so a later polygon overwrites an earlier one.


// 16 background pixels of map row `my`, columns 16 * bx .. 16 * bx + 15 (both relative to CAR_MAP_ORIGIN, inside the window):
// words [w0, w0 + nw) of the row.  chk = [2][2048] bytes, 0xFF where the column (axis 0) / row (axis 1) lies in a checker square.
__device__ __forceinline__ uint32_t road_bg_word(const uint8_t* chk, int bx, int my, int w, uint32_t grass4, uint32_t check4) {
    const uint32_t fx = __ldg(reinterpret_cast<const uint32_t*>(chk) + bx * 4 + w);
    const uint32_t fy = __ldg(chk + 2048 + my) ? 0xFFFFFFFFu : 0u;
    const uint32_t f = fx & fy;
    return (f & check4) | (~f & grass4);
}

__device__ inline void paint_polygon_warp(uint16_t* index, uint8_t* blocks, int& n_blocks, const short* vx, const short* vy, int n,
                                          uint8_t gray, int lane, int32_t* dropped, const uint8_t* chk, uint32_t grass4, uint32_t check4) {
    int miny = vy[0], maxy = vy[0];
    for (int i = 1; i < n; ++i) { miny = min(miny, (int)vy[i]); maxy = max(maxy, (int)vy[i]); }
    constexpr int HI = CAR_MAP_ORIGIN + CAR_MAP_GRID * CAR_MAP_BLOCK;
    for (int r0 = miny; r0 <= maxy; r0 += 32) {
        const int V = r0 + lane;
        const bool active = V <= maxy;
        short4 sp = make_short4(1, 0, 1, 0);
        if (active) sp = scanline_spans(vx, vy, n, V, maxy);
        const int by = (V - CAR_MAP_ORIGIN) >> 4;
        const bool row_ok = active && by >= 0 && by < CAR_MAP_GRID;
        if (active && !row_ok && (sp.x <= sp.y || sp.z <= sp.w)) atomicAdd(dropped, 1);
        // x ranges clipped to the window; pixels outside it are dropped (flagged)
        int x0[2] = {sp.x, sp.z}, x1[2] = {sp.y, sp.w};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!row_ok) { x0[k] = 1; x1[k] = 0; continue; }
            if (x0[k] <= x1[k] && (x0[k] < CAR_MAP_ORIGIN || x1[k] >= HI)) {
                atomicAdd(dropped, 1);
                x0[k] = max(x0[k], CAR_MAP_ORIGIN); x1[k] = min(x1[k], HI - 1);
            }
        }
        // ---- blocks this row needs that do not exist yet: allocated one at a time by the whole warp ----
        for (;;) {
            int need = -1;
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (need < 0 && x0[k] <= x1[k])
                    for (int bx = (x0[k] - CAR_MAP_ORIGIN) >> 4; bx <= ((x1[k] - CAR_MAP_ORIGIN) >> 4); ++bx)
                        if (__ldcg(index + by * CAR_MAP_GRID + bx) == 0) { need = by * CAR_MAP_GRID + bx; break; }
            const unsigned m = __ballot_sync(0xffffffffu, need >= 0);
            if (m == 0u) break;
            const int cell = __shfl_sync(0xffffffffu, need, __ffs(m) - 1);
            if (n_blocks >= CAR_MAP_MAX_BLOCKS) {
                if (lane == 0) { __stcg(index + cell, (uint16_t)0xFFFFu); atomicAdd(dropped, 1); }
            } else {
                const int gx = cell & (CAR_MAP_GRID - 1), my = (cell / CAR_MAP_GRID) * 16 + (lane >> 1), w = (lane & 1) * 2;   // lane = half a row
                reinterpret_cast<uint2*>(blocks + (size_t)n_blocks * 256)[lane] =
                    make_uint2(road_bg_word(chk, gx, my, w, grass4, check4), road_bg_word(chk, gx, my, w + 1, grass4, check4));
                if (lane == 0) __stcg(index + cell, (uint16_t)(n_blocks + 1));
                n_blocks += 1;
            }
            __syncwarp();
        }
        // ---- paint ----
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (x0[k] > x1[k]) continue;
            const int ry = (V - CAR_MAP_ORIGIN) & 15;
            for (int bx = (x0[k] - CAR_MAP_ORIGIN) >> 4; bx <= ((x1[k] - CAR_MAP_ORIGIN) >> 4); ++bx) {
                const unsigned idx = __ldcg(index + by * CAR_MAP_GRID + bx);
                if (idx == 0xFFFFu) continue;
                uint8_t* row = blocks + (size_t)(idx - 1u) * 256 + ry * 16;
                const int xa = max(x0[k] - CAR_MAP_ORIGIN - bx * 16, 0), xb = min(x1[k] - CAR_MAP_ORIGIN - bx * 16, 15);
                for (int x = xa; x <= xb; ++x) row[x] = gray;
            }
        }
        __syncwarp();
    }
}

// Road map of the track in `slot` (n_track tiles), by one warp.
__device__ inline void paint_road_map(const CarDev& p, int slot, int n_track, int lane) {
    const CarTile* tiles = p.tiles + (size_t)slot * CAR_MAX_TRACK;
    uint16_t* index = p.map_index + (size_t)slot * CAR_MAP_GRID * CAR_MAP_GRID;
    uint8_t* blocks = p.map_blocks + (size_t)slot * CAR_MAP_MAX_BLOCKS * 256;
    const uint8_t* G = p.consts->gray;
    for (int i = lane; i < CAR_MAP_GRID * CAR_MAP_GRID / 8; i += 32) __stcg(reinterpret_cast<uint4*>(index) + i, make_uint4(0u, 0u, 0u, 0u));
    __syncwarp();
    int n_blocks = 0;
    const uint32_t grass4 = 0x01010101u * G[G_GRASS], check4 = 0x01010101u * G[G_CHECK];
    for (int t = n_track - 1; t >= 0; --t) {
        const CarTile* T = tiles + t;
        paint_polygon_warp(index, blocks, n_blocks, T->mx, T->my, 5, G[G_ROAD0 + t % 3], lane, p.overrun + 1, p.chk, grass4, check4);
        const uint8_t flags = T->flags;
        if (flags & 2)
            paint_polygon_warp(index, blocks, n_blocks, T->kmx, T->kmy, 4, (flags & 4) ? G[G_KERB_W] : G[G_KERB_R], lane, p.overrun + 1, p.chk,
                               grass4, check4);
    }
    __threadfence();
}

}  // namespace crl
