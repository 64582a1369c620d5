// car_raster.cu -- 96x96 grayscale observation of cCarRacing, one CTA per (env, player) frame.
//
// Replaces (paths relative to /root/reference/competitive_rl/): get_observation (car_racing_multi_players.py
// :622-634), camera_update (:791-812), camera_view (:764-789), render_road_for_observation_map (:732-755),
// render(mode="internal_rgb_array") (:857-863), Car.draw_for_pygame (car_dynamics.py:284-298),
// render_indicators_for_pygame (:645-670) + pygame_rendering.py:8-18, and the FrameStack /
// MultipleFrameStack + FlattenMultiAgentObservation + WrapPyTorch layout (utils/atari_wrappers.py:222-334).
//
// The reference paints the road once per reset into a 10 000 x 10 000 px surface (400 MB) at obs_scale
// px/unit, crops 192x192 around the camera, rotates the crop (pygame.transform.rotate, 16.16 fixed-point
// nearest neighbour) and centre-blits it to the 96x96 screen, then draws the cars' fixture polygons and the
// HUD.  Here the painted part of that surface is kept per track as a sparse raster of 16x16 px blocks
// (car_spans.cuh: painted once per track by the generator, off the step's critical path).  Per frame the CTA
// stages the blocks under the visible window into shared memory (<= 9 x 9 blocks: the 96 x 86 px above the HUD bar,
// rotated, span at most 129 map pixels per axis), pushes every screen pixel through the reference's integer
// pipeline (screen -> rotated surface -> source crop -> road-map pixel (U, V)) and reads its colour there
// (0 = background: grass / checker square by (U, V)); the <= 16 car fixture polygons (screen space, a few
// pixels each) are scan-converted with pygame 1.9's draw_fillpoly rule into a small span table, binned to
// 8x8-pixel cells and tested only in the cells they touch, the largest paint order winning.  pygame's integer
// rules are restated from memory exactly as in oracle/ref_shim/pygame, under which the reference's own
// renderer reproduces these frames bit for bit (tests/golden/car_frames.npz); parity against a real pygame
// build is unpinned.
#include <math.h>

#include "car_common.cuh"
#include "car_spans.cuh"

namespace crl {

#define CR_SIZE 0.02
#define CR_PLAYFIELD (2000.0 / 6.0)
#define CR_TRACK_WIDTH (40.0 / 6.0)
#define CR_BORDER (8.0 / 6.0)
#define CR_TRACK_DETAIL_STEP (21.0 / 6.0)

constexpr int RASTER_THREADS = 256;
constexpr int CAR_POLYS = 8 * CAR_MAX_PLAYERS;   // 4 wheels + 4 hull fixtures per car
constexpr int POLY_ROWS = 8;               // span rows kept per car polygon (a fixture is <= 5.2 px across at obs_scale: <= 7 rows)
constexpr int HUD_TOP = 86;                // (int)(H - 4 * (H / 40.0)) = (int)86.4: first row of the black HUD bar
// the pixels above the HUD bar as quads of 4 horizontally adjacent pixels, row-major: a warp walks 32 consecutive quads at a time
constexpr int QUADS_X = CAR_W / 4, WALK_QUADS = QUADS_X * HUD_TOP;
constexpr int HUD_WARP = 5;                // the warp that paints the HUD
// The 96 x 86 px above the HUD bar, rotated by any angle, cover at most floor(sqrt(95^2 + 85^2)) + 2 = 129 consecutive
// road-map columns / rows, i.e. at most 9 blocks of 16 per axis whatever the alignment (15 + 129 = 144).
constexpr int CROP_BLOCKS = 9, CROP_DIM = CROP_BLOCKS * CAR_MAP_BLOCK;

__constant__ float c_hull_poly[4][8][2] = {
    {{-60, +130}, {+60, +130}, {+60, +110}, {-60, +110}},
    {{-15, +120}, {+15, +120}, {+20, +20}, {-20, 20}},
    {{+25, +20}, {+50, -10}, {+50, -40}, {+20, -90}, {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20}},
    {{-50, -120}, {+50, -120}, {+50, -90}, {-50, -90}}};
__constant__ int c_hull_count[4] = {4, 4, 8, 4};

// Integer parameters of one frame's screen -> road-map mapping (pygame.transform.rotate + blit + subsurface), 64 bytes
struct FrameMap {
    int rx, ry;            // top-left of the 192x192 crop in the road map
    int obx, oby;          // first block (of the slot's block grid, may be < 0) of the staged window
    int bx, by;            // blit position of the rotated surface on the screen
    int nbx, nby_mul;      // blocks per axis of the staged window (<= CROP_BLOCKS): nbx, nby | mul << 8 with b / nbx = (b * mul) >> 10
    int isin, icos;
    int cx0, cy0;          // dx = cx0 + icos*x - isin*y ; dy = cy0 + isin*x + icos*y   (16.16)
    float camx, camy;      // camera_offset (b2Vec2)
    float ts, tc;          // sin/cos of tmp.angle = -camera_angle (fp32, b2Rot)
};
__device__ __forceinline__ double car_obs_scale() { return (10 / (100 / sqrt(96.0))) * 1.8; }   // CarRacing.obs_scale, :215

// Screen pixel (X, Y) -> road-map pixel: with (x, y) = (X - bx, Y - by), dx = cx0 + icos * x - isin * y and
// dy = cy0 + isin * x + icos * y in 16.16 fixed point, (U, V) = (rx + (dx >> 16), ry + (dy >> 16)) -- evaluated
// incrementally in walk_cells.  The visible 96x96 window is the centre of the rotated 192x192 crop, whose inscribed
// circle (radius 96) always contains it (half diagonal 68), so every screen pixel has a source inside the crop.

// pygame.draw.rect(screen, color, (x, y, w, h)) = polygon (l, t), (r, t), (r, b), (l, b), r = x + w - 1, b = y + h - 1
__device__ __forceinline__ void hud_rect(uint8_t* img, double x, double y, double w, double h, uint8_t val, int tid, int nthreads) {
    const int X = (int)x, Y = (int)y, W = (int)w, H = (int)h;
    const int l = X, r = X + W - 1, t = Y, b = Y + H - 1;
    if (t == b) return;                                   // every edge horizontal or degenerate: nothing is filled
    const int x0 = max(min(l, r), 0), x1 = min(max(l, r), CAR_W - 1), y0 = max(min(t, b), 0), y1 = min(max(t, b), CAR_H - 1);
    const int bw = x1 - x0 + 1, total = bw * (y1 - y0 + 1);
    if (bw <= 0 || total <= 0) return;
    for (int q = tid; q < total; q += nthreads) img[(y0 + q / bw) * CAR_W + x0 + q % bw] = val;
}

// one car polygon of the frame (screen pixels): rows [miny, miny + rows) of its span table, gray value
struct __align__(8) PolyMeta { short miny, rows; unsigned char gray, n, pad0, pad1; };

// What car_frame_setup_kernel hands to the render kernel besides the FrameMap, per frame (1632 bytes): the car polygons
// scan-converted with pygame's fill rule, the block-pool positions of the road-map blocks under the window, and the
// checker flags of the window's columns and rows (background of blocks nothing was painted in).
struct __align__(16) FrameAux {
    short4 spans[CAR_POLYS][POLY_ROWS];
    PolyMeta meta[CAR_POLYS];
    uint16_t blk[96];                              // [j * nbx + i]: CarDev::map_index entry of block (obx + i, oby + j); 0 = nothing painted
    uint8_t chkx[CROP_DIM], chky[CROP_DIM];        // 0xFF where column / row i of the staged window lies in a checker square (CarDev::chk)
};

struct RasterSmem {
    uint8_t img[CAR_PIX];
    uint8_t crop[CROP_DIM * CROP_DIM];             // the road map under the visible window, [v][u] from block (obx, oby)
    FrameAux aux;
    double hud_vals[8];
    int copy_next, hud_late;                       // ring -> terminal observation chunk counter; 1 = an indicator reaches above the bar
};

// Pixels of the frame above the HUD bar: lane = one quad (x .. x + 3, y), each pixel the road-map byte it samples from the
// staged window (background included) -- unless a car polygon was painted there before the walk (img starts out as
// CAR_UNPAINTED everywhere; no car pixel has that value).
constexpr uint32_t CAR_UNPAINTED = 0xFFu;
__device__ __forceinline__ void walk_quads(RasterSmem& S, const FrameMap& fm, int tid) {
    uint32_t* img32 = reinterpret_cast<uint32_t*>(S.img);
    // road-map pixel (rx + u, ry + v) sits at [v + offy][u + offx] of the staged window.  No bounds test: the 96x96 window
    // is the centre of the rotated 192x192 crop, whose inscribed circle (radius 96) contains it (half diagonal 68), so
    // every screen pixel samples inside the crop, and the staged blocks cover everything the rows above the HUD bar sample.
    const uint8_t* crop = S.crop + (fm.ry - CAR_MAP_ORIGIN - CAR_MAP_BLOCK * fm.oby) * CROP_DIM + (fm.rx - CAR_MAP_ORIGIN - CAR_MAP_BLOCK * fm.obx);
    const int bdx = fm.cx0 - fm.icos * fm.bx + fm.isin * fm.by, bdy = fm.cy0 - fm.isin * fm.bx - fm.icos * fm.by;
#pragma unroll 1
    for (int q = tid; q < WALK_QUADS; q += RASTER_THREADS) {
        const int y = (q * 2731) >> 16, x = (q - y * QUADS_X) * 4;     // q / 24 for q < 4096
        // dx = cx0 + icos * (x - bx) - isin * (y - by), dy = cy0 + isin * (x - bx) + icos * (y - by)   (16.16)
        int dx = bdx + fm.icos * x - fm.isin * y, dy = bdy + fm.isin * x + fm.icos * y;
        unsigned int k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned int u = __byte_perm((unsigned)dx, 0u, 0x4442), v = __byte_perm((unsigned)dy, 0u, 0x4442);   // (d >> 16) & 255: 0..191
            k[i] = crop[v * CROP_DIM + u];
            dx += fm.icos; dy += fm.isin;
        }
        uint32_t w = k[0] | (k[1] << 8) | (k[2] << 16) | (k[3] << 24);
        const uint32_t cars = img32[q];
        if (cars != 0x01010101u * CAR_UNPAINTED) {
            const uint32_t keep = __vcmpeq4(cars, 0x01010101u * CAR_UNPAINTED);      // 0xFF where no car pixel
            w = (w & keep) | (cars & ~keep);
        }
        img32[q] = w;
    }
}

// The car polygons (Car.draw_for_pygame, car_dynamics.py:284-298: for k in cars: wheels, then hull fixtures), by one warp,
// before the road pixels are walked.  The four wheels of a car share one colour and so do its four hull fixtures, so
// only the order of these layers matters: car 0 wheels < car 0 hull < car 1 wheels < car 1 hull.  lane = (polygon of the
// layer, row of its span table).
__device__ void paint_cars(RasterSmem& S, int players, int lane) {
    uint8_t* img = S.img;
    for (int layer = 0; layer < 2 * players; ++layer) {
        const int id = layer * 4 + (lane >> 3), r = lane & 7;
        const PolyMeta m = S.aux.meta[id];
        const int y = m.miny + r;
        if (r < m.rows && y >= 0 && y < HUD_TOP) {                    // rows under the HUD bar belong to the HUD
            const short4 sp = S.aux.spans[id][r];
            for (int x = max((int)sp.x, 0); x <= min((int)sp.y, CAR_W - 1); ++x) img[y * CAR_W + x] = m.gray;
            for (int x = max((int)sp.z, 0); x <= min((int)sp.w, CAR_W - 1); ++x) img[y * CAR_W + x] = m.gray;
        }
        __syncwarp();
    }
}

// Two kinds of passes run over a list of envs instead of over all of them: the auto-reset passes (only_done: the done
// list the post-step render built) and the pass over the envs of the slow physics pass (which == 2: the slow list of the
// sensor kernel).  Position i on the list -> frame; the render kernel of such a pass is launched with at most
// LIST_PASS_CTAS blocks that stride over the list.
constexpr int LIST_PASS_CTAS = 148 * 6;
__device__ __forceinline__ int listed_frames(const CarDev& p, int only_done, int which) {
    if (only_done) return min(*p.done_count, p.n) * p.players;
    if (which == 2) return min(*p.slow_count, p.n) * p.players;
    return p.n * p.players;
}
__device__ __forceinline__ int listed_frame(const CarDev& p, int only_done, int which, int i) {
    if (!only_done && which != 2) return i;
    const int k = i / p.players;
    return (only_done ? p.done_list[k] : p.slow_list[k]) * p.players + (i - k * p.players);
}

// Per-frame setup, part 1, one thread per (env, player) frame: camera and the integer screen -> road-map mapping (a serial
// fp64 chain that would stall a whole CTA of the render kernel).
__device__ FrameMap car_frame_map_of(const CarDev& p, int frame) {
    const CarHullConst* K = p.consts;
    const double obs_scale = car_obs_scale();
    FrameMap m;
    // ---- camera_update("rgb_array"): hull.position + R(angle) * (0, 16) ----
    const float4 b0 = *reinterpret_cast<const float4*>(p.body + (size_t)frame * 40);     // cx, cy, angle, vx
    const float bvy = p.body[(size_t)frame * 40 + 4];
    float hs, hc;
    sincosf(b0.z, &hs, &hc);
    const float hx = b0.x - (hc * K->hull_lcx - hs * K->hull_lcy), hy = b0.y - (hs * K->hull_lcx + hc * K->hull_lcy);
    double angle = (double)b0.z;
    const double vx = (double)b0.w, vy = (double)bvy;
    if (vx * vx + vy * vy > 0.5 * 0.5) angle = atan2(-vx, +vy);
    const float fa = (float)angle;
    float fs, fc;
    sincosf(fa, &fs, &fc);
    m.camx = hx + (fc * 0.0f - fs * 16.0f);
    m.camy = hy + (fs * 0.0f + fc * 16.0f);
    // ---- camera_view: crop rectangle, rotation, blit ----
    const double pos0 = obs_scale * -(double)m.camx + 5000.0, pos1 = obs_scale * -(double)m.camy + 5000.0;
    m.rx = (int)(pos0 - CAR_W); m.ry = (int)(pos1 - CAR_H);
    const int sw = 2 * CAR_W, sh = 2 * CAR_H;
    const double rad = (57.295779513 * angle) * .01745329251994329;
    const double sangle = sin(rad), cangle = cos(rad);
    const double cx = cangle * sw, cy = cangle * sh, sx = sangle * sw, sy = sangle * sh;
    const int nx = (int)fmax(fmax(fmax(fabs(cx + sy), fabs(cx - sy)), fabs(-cx + sy)), fabs(-cx - sy));
    const int ny = (int)fmax(fmax(fmax(fabs(sx + cy), fabs(sx - cy)), fabs(-sx + cy)), fabs(-sx - cy));
    const int cyi = ny / 2;
    const int xd = (sw - nx) * 32768, yd = (sh - ny) * 32768;
    m.isin = (int)(sangle * 65536); m.icos = (int)(cangle * 65536);
    const int ax = (nx * 32768) - (int)(cangle * (double)((nx - 1) * 32768));
    const int ay = (ny * 32768) - (int)(sangle * (double)((nx - 1) * 32768));
    m.cx0 = ax + xd + m.isin * cyi;
    m.cy0 = ay + yd - m.icos * cyi;
    m.bx = -(nx >> 1) + CAR_W / 2; m.by = -(ny >> 1) + CAR_H / 2;
    const float ta = (float)(-angle);
    sincosf(ta, &m.ts, &m.tc);
    // ---- road-map columns / rows the pixels above the HUD bar sample: the mapping is affine before the floor, so the
    //      extremes are at the corners of the 96 x 86 rectangle ----
    int u0 = 0x7fffffff, u1 = -0x7fffffff, v0 = 0x7fffffff, v1 = -0x7fffffff;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int x = ((c & 1) ? CAR_W - 1 : 0) - m.bx, y = ((c & 2) ? HUD_TOP - 1 : 0) - m.by;
        const int u = ((m.cx0 + m.icos * x - m.isin * y) >> 16) & 255, v = ((m.cy0 + m.isin * x + m.icos * y) >> 16) & 255;
        u0 = min(u0, u); u1 = max(u1, u); v0 = min(v0, v); v1 = max(v1, v);
    }
    m.obx = (m.rx + u0 - CAR_MAP_ORIGIN) >> 4; m.oby = (m.ry + v0 - CAR_MAP_ORIGIN) >> 4;
    m.nbx = min(((m.rx + u1 - CAR_MAP_ORIGIN) >> 4) - m.obx + 1, CROP_BLOCKS);
    const int nby = min(((m.ry + v1 - CAR_MAP_ORIGIN) >> 4) - m.oby + 1, CROP_BLOCKS);
    m.nby_mul = nby | (((1024 + m.nbx - 1) / m.nbx) << 8);
    return m;
}

__global__ void __launch_bounds__(128)
car_frame_setup_kernel(CarDev p, int only_done, int which) {
    int frame = blockIdx.x * blockDim.x + threadIdx.x;             // env * players + player; auto-reset pass: position on the done list
    if (frame >= listed_frames(p, only_done, which)) return;
    frame = listed_frame(p, only_done, which, frame);
    const int e = frame / p.players;
    if (which == 1 && p.deferred[e] != 0) return;                  // the envs of the slow physics pass come in their own pass (over the slow list)
    p.frame_map[frame] = car_frame_map_of(p, frame);
}

// Car.draw_for_pygame (car_dynamics.py:284-298) for one fixture polygon: `part` 0..3 = the wheels, 4..7 = the hull fixtures
// of the car whose bodies start at `b`.  FixturePose = where the fixture's body sits; fixture_vertex = vertex i pushed through
// path = -scale * (tmp * ((trans * v) - offset)) + (W/2, H/2) in b2Vec2 fp32 arithmetic and truncated to int like pygame
// (clamped for the short storage: a polygon that far off the screen cannot touch the window anyway).
struct FixturePose { float bs, bc, px, py; int n; };
__device__ __forceinline__ FixturePose fixture_pose(const CarHullConst* K, const float* b, int part) {
    const float* body = (part < 4) ? b + 8 * (part + 1) : b;
    FixturePose f;
    sincosf(body[2], &f.bs, &f.bc);
    f.px = body[0]; f.py = body[1];
    if (part >= 4) { f.px = b[0] - (f.bc * K->hull_lcx - f.bs * K->hull_lcy); f.py = b[1] - (f.bs * K->hull_lcx + f.bc * K->hull_lcy); }
    f.n = (part < 4) ? 4 : c_hull_count[part - 4];
    return f;
}
__device__ __forceinline__ void fixture_vertex(const FixturePose& f, const FrameMap& m, int part, int i, int& ix, int& iy) {
    const float hw = (float)(14 * CR_SIZE), hr = (float)(27 * CR_SIZE);
    const double obs_scale = car_obs_scale();
    float lx, ly;
    if (part < 4) { lx = (i == 0 || i == 1) ? hw : -hw; ly = (i == 1 || i == 2) ? hr : -hr; }
    else { lx = (float)(c_hull_poly[part - 4][i][0] * CR_SIZE); ly = (float)(c_hull_poly[part - 4][i][1] * CR_SIZE); }
    const float wx = (f.bc * lx - f.bs * ly) + f.px, wy = (f.bs * lx + f.bc * ly) + f.py;
    const float ox = wx - m.camx, oy = wy - m.camy;
    const float rx2 = (m.tc * ox - m.ts * oy) + 0.0f, ry2 = (m.ts * ox + m.tc * oy) + 0.0f;
    const float sxp = (float)((double)rx2 * -obs_scale) + (float)(CAR_W / 2.0);
    const float syp = (float)((double)ry2 * -obs_scale) + (float)(CAR_H / 2.0);
    ix = max(-32000, min(32000, (int)sxp)); iy = max(-32000, min(32000, (int)syp));
}

// Per-frame setup, part 2, 16 threads per frame: one car polygon each -- b2Vec2 fp32 arithmetic: path = -scale * (tmp *
// ((trans * v) - offset)) + (W/2, H/2), truncated to int by pygame -- scan-converted with draw_fillpoly's rule, and the
// pool positions of the road-map blocks under the window.
__global__ void __launch_bounds__(128, 8)
car_frame_aux_kernel(CarDev p, int only_done, int which) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    int frame = gt >> 4;
    const int l = gt & 15;
    if (frame >= listed_frames(p, only_done, which)) return;
    frame = listed_frame(p, only_done, which, frame);
    const int e = frame / p.players, pi = frame - e * p.players;
    if (which == 1 && p.deferred[e] != 0) return;                  // the envs of the slow physics pass come in their own pass (over the slow list)
    const CarHullConst* K = p.consts;
    const FrameMap& m = p.frame_map[frame];
    FrameAux* aux = reinterpret_cast<FrameAux*>(p.frame_aux) + frame;
    // ---- the road-map blocks under the window ----
    {
        const int slot = car_slot(p, e);
        const uint16_t* index = p.map_index + (size_t)slot * CAR_MAP_GRID * CAR_MAP_GRID;
        const int nb = m.nbx * (m.nby_mul & 255), mul = m.nby_mul >> 8;
        for (int bq = l; bq < nb; bq += 16) {
            const int j = (bq * mul) >> 10, i = bq - j * m.nbx;
            const int gx = m.obx + i, gy = m.oby + j;
            unsigned int idx = 0u;
            if ((unsigned)gx < (unsigned)CAR_MAP_GRID && (unsigned)gy < (unsigned)CAR_MAP_GRID) idx = index[gy * CAR_MAP_GRID + gx];
            aux->blk[bq] = (uint16_t)(idx == 0xFFFFu ? 0u : idx);   // a dropped block (flagged when painted) shows the background
        }
    }
    // ---- checker flags of the window's columns and rows: 16-byte pieces, lanes 0..8 the columns, the rest + a second round the rows ----
    for (int t = l; t < 2 * CROP_BLOCKS; t += 16) {
        const int axis = t >= CROP_BLOCKS, k = axis ? t - CROP_BLOCKS : t, g = (axis ? m.oby : m.obx) + k;
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)g < (unsigned)CAR_MAP_GRID) f = __ldg(reinterpret_cast<const uint4*>(p.chk + axis * 2048) + g);
        *reinterpret_cast<uint4*>((axis ? aux->chky : aux->chkx) + k * 16) = f;
    }
    // ---- car polygon l = car * 8 + part: parts 0..3 the wheels, 4..7 the hull fixtures ----
    PolyMeta pm;
    pm.miny = 0; pm.rows = 0; pm.gray = 0; pm.n = 0; pm.pad0 = pm.pad1 = 0;
    const int ck = l >> 3, part = l & 7;
    if (ck < p.players) {
        const FixturePose fp = fixture_pose(K, p.body + ((size_t)e * p.players + ck) * 40, part);
        const int n = fp.n;
        short vx[8], vy[8];
        int minx = 0x7fffffff, maxx = -0x7fffffff, miny = 0x7fffffff, maxy = -0x7fffffff;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            vx[i] = 0; vy[i] = 0;
            if (i < n) {
                int ix, iy;
                fixture_vertex(fp, m, part, i, ix, iy);
                vx[i] = (short)ix; vy[i] = (short)iy;
                minx = min(minx, ix); maxx = max(maxx, ix); miny = min(miny, iy); maxy = max(maxy, iy);
            }
        }
        pm.gray = (part < 4) ? K->gray[G_WHEEL] : ((ck == pi) ? K->gray[G_OWN] : K->gray[G_OTHER]);
        pm.n = (unsigned char)n;
        pm.miny = (short)miny;
        if (maxx >= 0 && minx < CAR_W && maxy >= 0 && miny < HUD_TOP) {         // can touch the rows above the HUD bar
            int rows = maxy - miny + 1;
            // a fixture is at most 5.3 px across at obs_scale (rigid polygons), i.e. <= 7 rows; anything taller is cut (flagged)
            if (rows > POLY_ROWS) { atomicAdd(p.overrun + 2, 1); rows = POLY_ROWS; }
            pm.rows = (short)rows;
            for (int r = 0; r < rows; ++r) aux->spans[l][r] = scanline_spans(vx, vy, n, miny + r, maxy);
        }
    }
    aux->meta[l] = pm;
}

// The same for passes over few frames (the auto-reset and slow-list passes, small batches), where what counts is the
// length of one thread's chain, not the instruction total: 128 threads per frame, thread (polygon, row) -- lane i of a
// polygon's eight computes vertex i, the eight exchange them by shuffles, and each scan-converts ONE row.  Twice the
// instructions per frame, a third of the latency.
constexpr int AUX_ROW_THREADS = 128;
__global__ void __launch_bounds__(2 * AUX_ROW_THREADS)
car_frame_aux_rows_kernel(CarDev p, int only_done, int which) {
    int frame = blockIdx.x * 2 + (threadIdx.x >> 7);
    const int t = threadIdx.x & (AUX_ROW_THREADS - 1);
    if (frame >= listed_frames(p, only_done, which)) return;
    frame = listed_frame(p, only_done, which, frame);
    const int e = frame / p.players, pi = frame - e * p.players;
    if (which == 1 && p.deferred[e] != 0) return;
    const CarHullConst* K = p.consts;
    // part 1 (camera, screen -> road-map mapping) by the first thread of the frame's 128: no separate launch for these passes
    __shared__ FrameMap s_fm[2];
    if (t == 0) { s_fm[threadIdx.x >> 7] = car_frame_map_of(p, frame); p.frame_map[frame] = s_fm[threadIdx.x >> 7]; }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 7)), "r"(AUX_ROW_THREADS));   // the frame's four warps only
    const FrameMap& m = s_fm[threadIdx.x >> 7];
    FrameAux* aux = reinterpret_cast<FrameAux*>(p.frame_aux) + frame;
    {   // the road-map blocks under the window
        const int nb = m.nbx * (m.nby_mul & 255), mul = m.nby_mul >> 8;
        if (t < nb) {
            const int slot = car_slot(p, e);
            const uint16_t* index = p.map_index + (size_t)slot * CAR_MAP_GRID * CAR_MAP_GRID;
            const int j = (t * mul) >> 10, i = t - j * m.nbx;
            const int gx = m.obx + i, gy = m.oby + j;
            unsigned int idx = 0u;
            if ((unsigned)gx < (unsigned)CAR_MAP_GRID && (unsigned)gy < (unsigned)CAR_MAP_GRID) idx = index[gy * CAR_MAP_GRID + gx];
            aux->blk[t] = (uint16_t)(idx == 0xFFFFu ? 0u : idx);
        }
    }
    if (t >= AUX_ROW_THREADS - 2 * CROP_BLOCKS) {   // checker flags of the window's columns and rows, in 16-byte pieces
        const int q = t - (AUX_ROW_THREADS - 2 * CROP_BLOCKS);
        const int axis = q >= CROP_BLOCKS, k = axis ? q - CROP_BLOCKS : q, g = (axis ? m.oby : m.obx) + k;
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)g < (unsigned)CAR_MAP_GRID) f = __ldg(reinterpret_cast<const uint4*>(p.chk + axis * 2048) + g);
        *reinterpret_cast<uint4*>((axis ? aux->chky : aux->chkx) + k * 16) = f;
    }
    // car polygon id = car * 8 + part (parts 0..3 the wheels, 4..7 the hull fixtures), row r of its span table
    const int id = t >> 3, r = t & 7;
    const int ck = id >> 3, part = id & 7;
    PolyMeta pm;
    pm.miny = 0; pm.rows = 0; pm.gray = 0; pm.n = 0; pm.pad0 = pm.pad1 = 0;
    if (ck < p.players) {                                           // uniform over the warp (four polygons of one car)
        const FixturePose fp = fixture_pose(K, p.body + ((size_t)e * p.players + ck) * 40, part);
        const int n = fp.n;
        int ix = 0, iy = 0;
        if (r < n) fixture_vertex(fp, m, part, r, ix, iy);          // vertex r of the polygon
        short vx[8], vy[8];
        int minx = 0x7fffffff, maxx = -0x7fffffff, miny = 0x7fffffff, maxy = -0x7fffffff;
        const int lane0 = threadIdx.x & 24;                         // first lane of this polygon's eight within the warp
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gx = __shfl_sync(0xffffffffu, ix, lane0 + i), gy = __shfl_sync(0xffffffffu, iy, lane0 + i);
            vx[i] = (short)gx; vy[i] = (short)gy;
            if (i < n) { minx = min(minx, gx); maxx = max(maxx, gx); miny = min(miny, gy); maxy = max(maxy, gy); }
        }
        pm.gray = (part < 4) ? K->gray[G_WHEEL] : ((ck == pi) ? K->gray[G_OWN] : K->gray[G_OTHER]);
        pm.n = (unsigned char)n;
        pm.miny = (short)miny;
        if (maxx >= 0 && minx < CAR_W && maxy >= 0 && miny < HUD_TOP) {
            int rows = maxy - miny + 1;
            if (rows > POLY_ROWS) { if (r == 0) atomicAdd(p.overrun + 2, 1); rows = POLY_ROWS; }
            pm.rows = (short)rows;
            if (r < rows) aux->spans[id][r] = scanline_spans(vx, vy, n, miny + r, maxy);
        }
    }
    if (r == 0) aux->meta[id] = pm;
}

// HUD indicators and reward text (render_indicators_for_pygame :645-670), one warp, in paint order
__device__ void paint_hud_indicators(const RasterSmem& S, const uint8_t* glyphs, const uint8_t* G, uint8_t* img, int lane) {
    const double W = CAR_W, H = CAR_H, s = W / 40.0, h = H / 40.0;
    hud_rect(img, 5 * s, H - h, s, h * (-0.02 * S.hud_vals[0]), G[G_BLUE], lane, 32);
    __syncwarp();
    for (int k = 0; k < 4; ++k) {
        hud_rect(img, (7 + k) * s, H - h, s, h * (-0.01 * S.hud_vals[1 + k]), k < 2 ? G[G_BLUE] : G[G_BLUE2], lane, 32);
        __syncwarp();
    }
    hud_rect(img, 20 * s, H - 2 * h, s * (10.0 * S.hud_vals[5]), 2 * h, G[G_GREEN], lane, 32);
    __syncwarp();
    hud_rect(img, 30 * s, H - 2 * h, s * (0.8 * S.hud_vals[6]), 2 * h, G[G_RED], lane, 32);
    __syncwarp();
    if (glyphs != nullptr) {     // draw_text("%05.0f" % reward) at (W/100, H - H/20): lane = one pixel of the 4x8 glyph cell
        // "%05.0f": round half to even, sign kept for negative values, zero padded to width 5
        const double rv = S.hud_vals[7];
        double mag = rint(fabs(rv));
        char digits[24];
        int nd = 0;
        if (mag == 0) digits[nd++] = 0;
        {   // integer digits (exact, no fp64 divisions); a reward is bounded by 1000 + 0.1 per step, far below 2^32
            unsigned int mi = (unsigned int)fmin(mag, 4.0e9);
            while (mi != 0u) { const unsigned int qi = mi / 10u; digits[nd++] = (char)(mi - qi * 10u); mi = qi; }
        }
        const int neg = (rv < 0 || (rv == 0 && signbit(rv))) ? 1 : 0;
        const int body = nd + neg;
        int pen = (int)(W / 100);
        const int y0 = (int)(H - H / 20);
        const int pad = 5 > body ? 5 - body : 0;
        const int gx = lane & 3, gy = lane >> 2;
        for (int i = 0; i < body + pad; ++i) {
            int gi;
            if (neg && i == 0) gi = 10;
            else if (i < neg + pad) gi = 0;
            else gi = digits[nd - 1 - (i - neg - pad)];
            if (glyphs[(gi * 8 + gy) * 4 + gx]) {
                const int px = pen + gx, py = y0 + gy;
                if (px >= 0 && px < CAR_W && py >= 0 && py < CAR_H) img[py * CAR_W + px] = G[G_TEXT];
            }
            pen += glyphs[11 * 8 * 4 + gi];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(RASTER_THREADS, 7)
car_render_kernel(CarDev p, int only_done, int which, uint8_t* __restrict__ obs, uint8_t* __restrict__ term_obs) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RasterSmem& S = *reinterpret_cast<RasterSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_listed = listed_frames(p, only_done, which);
  for (int it = blockIdx.x; it < n_listed; it += gridDim.x) {        // a post-step pass has one CTA per frame: a single round
    const int frame = listed_frame(p, only_done, which, it);   // env * players + player
    const int e = (p.players == 2) ? frame >> 1 : frame;          // players is 1 or 2
    if (which == 1 && p.deferred[e] != 0) continue;              // the envs of the slow physics pass come in their own pass (over the slow list)
    if (tid == 0 && (p.players == 1 || (frame & 1) == 0)) {
        if (only_done) p.ring_pos[e] = p.c - 1;        // every ring slot holds the reset frame (nobody reads ring_pos in this pass)
        else if (p.collect_done && p.env_done[e]) {    // finished in this step: onto the list the auto-reset passes run over
            const int k = atomicAdd(p.done_count, 1);
            if (k < p.n) p.done_list[k] = e;
        }
    }
    const CarHullConst* K = p.consts;
    const uint8_t* G = K->gray;
    const FrameMap fm = p.frame_map[frame];           // written by car_frame_setup_kernel
    const int slot = car_slot(p, e);
    const int C = p.c;
    // stack mode: the frames live in an internal ring [frame][C]; the C - 1 that stay in the observation were moved
    // ring -> observation by car_stack_shift_kernel, the new one goes to its ring slot and to channel C - 1.  Ring mode: the
    // observation buffer itself is a double-write ring of 2C slots (the new frame goes to slots k and k + C, the caller
    // looks at slots k+1 .. k+C), so nothing is ever moved.
    // Output layout: [env][players * C][96][96]: player-major channel blocks (FlattenMultiAgentObservation concatenates
    // the players' stacks on the channel axis), oldest frame first within a player.
    // Ahead-write stack mode (a rotation of registered buffers): the new frame goes to channel C-1 of the current buffer,
    // C-2 of the next, ... 0 of the C-1-th next; the current buffer already holds its older frames.
    const bool ringm = p.ring_mode != 0, ahead = p.rot_n > 0;
    uint8_t* ring = (ringm || ahead) ? nullptr : p.ring + (size_t)frame * C * CAR_PIX;
    const bool fill_all = only_done != 0 || ((ringm || ahead) ? p.fill_all != 0 : p.ring_pos[e] < 0);
    int newest = C - 1;
    if (ringm) newest = p.ring_phase;
    else if (!ahead && !fill_all) { newest = p.ring_pos[e] + 1; if (newest >= C) newest = 0; }
    uint8_t* out = (ahead ? p.rot[p.rot_pos] : obs) + (size_t)frame * (ringm ? 2 * C : C) * CAR_PIX;
    uint8_t* tout = (term_obs != nullptr && !only_done && p.env_done[e]) ? term_obs + (size_t)frame * C * CAR_PIX : nullptr;

    // what the setup kernel prepared: car polygon span tables and the block list (84 x 16 bytes)
    if (tid < (int)(sizeof(FrameAux) / 16))
        reinterpret_cast<uint4*>(&S.aux)[tid] = reinterpret_cast<const uint4*>(reinterpret_cast<const FrameAux*>(p.frame_aux) + frame)[tid];
    {   // every pixel starts out "no car here"; rows 86.. are the HUD warp's
        const uint32_t f4 = 0x01010101u * CAR_UNPAINTED;
        for (int q = tid; q < HUD_TOP * CAR_W / 16; q += RASTER_THREADS) reinterpret_cast<uint4*>(S.img)[q] = make_uint4(f4, f4, f4, f4);
    }
    if (tid == 255) { S.copy_next = 0; S.hud_late = 0; }
    if (tid >= 240 && tid < 248) {      // HUD inputs (render_indicators_for_pygame :645-670)
        const int k = tid - 240;
        const float* b = p.body + (size_t)frame * 40;
        const double* wd = p.wheel + (size_t)frame * 8;
        double v;
        if (k == 0) { const double vx = (double)b[3], vy = (double)b[4]; v = sqrt(vx * vx + vy * vy); }
        else if (k <= 4) v = wd[k - 1];
        else if (k == 5) v = (double)(b[8 + 2] - b[2]);     // wheels[0].joint.angle
        else if (k == 6) v = (double)b[5];                  // hull.angularVelocity
        else v = p.reward[2 * (size_t)frame];
        S.hud_vals[k] = v;
    }
    __syncthreads();
    // ---- three jobs side by side.  Warp 0 paints the car polygons (the walk below keeps them); warp HUD_WARP the HUD bar --
    //      rows 86..95, which the walk never writes; the six other warps stage the road map under the window: 16 threads per
    //      16 x 16 block, one 16-byte row each; where nothing was painted (no block, or outside the slot's grid) the row is
    //      the background: grass / checker squares ----
    if (warp == 0) paint_cars(S, p.players, lane);
    else if (warp != HUD_WARP) {
        const uint4* blocks = reinterpret_cast<const uint4*>(p.map_blocks + (size_t)slot * CAR_MAP_MAX_BLOCKS * 256);
        const int st = tid - 32 - (warp > HUD_WARP ? 32 : 0);       // 0..191 over warps 1-4, 6, 7
        const int nb = fm.nbx * (fm.nby_mul & 255), sub = st & 15, mul = fm.nby_mul >> 8;
        const uint32_t grass4 = 0x01010101u * G[G_GRASS], check4 = 0x01010101u * G[G_CHECK];
#pragma unroll 2
        for (int bq = st >> 4; bq < nb; bq += 12) {
            const int j = (bq * mul) >> 10, i = bq - j * fm.nbx;
            const unsigned int idx = S.aux.blk[bq];
            uint4 v;
            if (idx != 0u) v = blocks[(idx - 1u) * 16 + sub];
            else {
                uint4 f = *reinterpret_cast<const uint4*>(S.aux.chkx + i * 16);
                const uint32_t fy = S.aux.chky[j * 16 + sub] ? 0xFFFFFFFFu : 0u;
                f.x &= fy; f.y &= fy; f.z &= fy; f.w &= fy;
                v = make_uint4((f.x & check4) | (~f.x & grass4), (f.y & check4) | (~f.y & grass4), (f.z & check4) | (~f.z & grass4),
                               (f.w & check4) | (~f.w & grass4));
            }
            *reinterpret_cast<uint4*>(S.crop + (j * 16 + sub) * CROP_DIM + i * 16) = v;
        }
    }
    uint8_t* img = S.img;
    __syncwarp();
    if (warp == HUD_WARP) {
        // the indicators are painted after the scene in the reference (render_indicators_for_pygame :645-670); they stay inside
        // the bar unless a vertical one is taller than 7 px (speed >= 146, wheel omega >= 292): then they wait for the scene
        const double H = CAR_H, h = H / 40.0;
        bool late = false;
        for (int k = 0; k < 5; ++k) {
            const int Y = (int)(H - h), Hh = (int)(h * ((k == 0 ? -0.02 : -0.01) * S.hud_vals[k]));
            late = late || min(Y, Y + Hh - 1) < HUD_TOP;
        }
        const uint32_t g4 = 0x01010101u * (uint32_t)G[G_HUD];
        for (int q = lane; q < (CAR_H - HUD_TOP) * CAR_W / 16; q += 32) reinterpret_cast<uint4*>(img + HUD_TOP * CAR_W)[q] = make_uint4(g4, g4, g4, g4);
        __syncwarp();
        if (late) { if (lane == 0) S.hud_late = 1; }
        else paint_hud_indicators(S, p.glyphs, G, img, lane);
        __syncwarp();
    }
    // The C - 1 older frames were moved ring -> observation by car_stack_shift_kernel before this launch (stack mode; a
    // ring-mode observation needs no move at all).  Only the terminal observation of a finished env is assembled here.
    if (!fill_all && tout != nullptr) {
        constexpr int PER_SLOT = CAR_PIX / 16 / 64;                // chunks of 64 uint4 (two per lane) per frame: 9
        const int n_chunks = (C - 1) * PER_SLOT;
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&S.copy_next, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= n_chunks) break;
            const int sl = ch / PER_SLOT, q = (ch % PER_SLOT) * 64 + lane;       // slot sl of the output = the sl-th oldest frame
            int rs = newest + 1 + sl;                             // < 2 C
            if (!ringm && rs >= C) rs -= C;
            if (ahead) rs = sl;                                   // the current buffer already holds the frames that stay
            const uint4* rsrc = reinterpret_cast<const uint4*>(((ringm || ahead) ? out : ring) + (size_t)rs * CAR_PIX);
            const uint4 v0 = rsrc[q], v1 = rsrc[q + 32];
            uint4* tdst = reinterpret_cast<uint4*>(tout + (size_t)sl * CAR_PIX);
            tdst[q] = v0; tdst[q + 32] = v1;
        }
    }
    __syncthreads();
    // ---- pixels above the HUD bar ----
    walk_quads(S, fm, tid);
    __syncthreads();
    if (S.hud_late) {                                              // uniform over the CTA
        if (warp == HUD_WARP) paint_hud_indicators(S, p.glyphs, G, img, lane);
        __syncthreads();
    }
    const uint4* src = reinterpret_cast<const uint4*>(img);
    if (ahead) {
        uint4* tdst = (tout && !fill_all) ? reinterpret_cast<uint4*>(tout + (size_t)(C - 1) * CAR_PIX) : nullptr;
        for (int j = 0; j < C; ++j) {                          // buffer j calls ahead: this frame is its channel C-1-j
            int bi = p.rot_pos + j;
            if (bi >= p.rot_n) bi -= p.rot_n;
            uint8_t* buf = p.rot[bi] + (size_t)frame * C * CAR_PIX;
            // after a reset every older slot holds the reset frame too: channels 0 .. C-1-j of that buffer
            for (int ch = fill_all ? 0 : C - 1 - j; ch <= C - 1 - j; ++ch) {
                uint4* dst = reinterpret_cast<uint4*>(buf + (size_t)ch * CAR_PIX);
                for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) dst[q] = src[q];
            }
        }
        if (tdst)
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) tdst[q] = src[q];
    } else if (fill_all && ringm) {
        for (int sl = 0; sl < 2 * C; ++sl) {
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) dst[q] = src[q];
        }
    } else if (fill_all) {
        for (int sl = 0; sl < C; ++sl) {
            uint4* rdst = reinterpret_cast<uint4*>(ring + (size_t)sl * CAR_PIX);
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)sl * CAR_PIX);
            for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) { const uint4 vv = src[q]; rdst[q] = vv; dst[q] = vv; }
        }
    } else {
        // stack: internal ring slot + newest channel.  ring: slots k and k + C of the observation ring
        uint4* rdst = reinterpret_cast<uint4*>((ringm ? out : ring) + (size_t)newest * CAR_PIX);
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(ringm ? newest + C : C - 1) * CAR_PIX);
        uint4* tdst = tout ? reinterpret_cast<uint4*>(tout + (size_t)(C - 1) * CAR_PIX) : nullptr;
        for (int q = tid; q < CAR_PIX / 16; q += RASTER_THREADS) {
            const uint4 vv = src[q];
            rdst[q] = vv;
            dst[q] = vv;
            if (tdst) tdst[q] = vv;
        }
    }
    __syncthreads();                                   // the shared-memory frame is reused by the next round
  }
}

// FrameStack in stack mode (atari_wrappers.py:222-259): the observation is the C newest frames, oldest first.  The C - 1
// frames that stay are already known before the step: this kernel moves them from the internal ring to channels
// 0 .. C-2 of every (env, player), and the render kernel then only writes the new frame (ring slot + channel C-1).  It is
// pure DRAM traffic, and it neither reads what the render passes of the step write (the ring slot of the new frame is the
// one slot it does not read) nor writes what they write, so crl_car_step runs it on a side stream NEXT TO the main render
// pass, which is bound by instruction issue: a small persistent grid (two 192-thread blocks per SM, six 16-byte loads
// in flight per thread) that leaves the registers and shared memory of the SMs to the render CTAs.  Envs that were just
// reset (ring_pos < 0) are skipped: their first render fills every channel.
constexpr int SHIFT_THREADS = 192, SHIFT_FRAME_V = CAR_PIX / 16;    // 576 uint4 per frame = 3 per thread
__global__ void __launch_bounds__(SHIFT_THREADS) car_stack_shift_kernel(CarDev p, uint8_t* __restrict__ obs) {
    const int C = p.c, n_units = p.n * p.players * (C - 1);         // unit = one frame that stays: (env, player, channel)
    const uint4* ring = reinterpret_cast<const uint4*>(p.ring);
    uint4* out = reinterpret_cast<uint4*>(obs);
    for (int u0 = blockIdx.x; u0 < n_units; u0 += 2 * gridDim.x) {
        uint4 v[2][3];
        size_t dst[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int u = u0 + k * gridDim.x;
            dst[k] = (size_t)-1;
            if (u < n_units) {
                const int frame = u / (C - 1), sl = u - frame * (C - 1);
                const int pos = p.ring_pos[p.players == 2 ? frame >> 1 : frame];
                if (pos >= 0) {
                    int rs = pos + 2 + sl;                          // the slot after the one the new frame will take = the oldest that stays
                    rs -= (rs >= C) ? C : 0;
                    const uint4* src = ring + ((size_t)frame * C + rs) * SHIFT_FRAME_V + threadIdx.x;
                    v[k][0] = __ldcs(src); v[k][1] = __ldcs(src + SHIFT_THREADS); v[k][2] = __ldcs(src + 2 * SHIFT_THREADS);
                    dst[k] = ((size_t)frame * C + sl) * SHIFT_FRAME_V + threadIdx.x;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (dst[k] != (size_t)-1) {
                __stcs(out + dst[k], v[k][0]); __stcs(out + dst[k] + SHIFT_THREADS, v[k][1]); __stcs(out + dst[k] + 2 * SHIFT_THREADS, v[k][2]);
            }
    }
}

__global__ void car_ring_advance_kernel(CarDev p) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    p.ring_pos[e] = p.ring_pos[e] < 0 ? p.c - 1 : (p.ring_pos[e] + 1) % p.c;
}

// road-map bounds of the checker squares: for even grid cells g = -20 + 2i the square polygon
// (k*g + k, .), (k*g, .) ... is int-truncated by pygame; per axis [lo_i, hi_i] inclusive (host, fp64)
void car_checker_table(int* out /* [2][20][2] */) {
    const double k = CR_PLAYFIELD / 20.0, osc = (10 / (100 / sqrt(96.0))) * 1.8;
    for (int axis = 0; axis < 2; ++axis)
        for (int i = 0; i < 20; ++i) {
            const int g = -20 + 2 * i;
            const int a = (int)(osc * -(k * g + k) + 5000.0), b = (int)(osc * -(k * g + 0) + 5000.0);
            out[(axis * 20 + i) * 2] = a < b ? a : b;
            out[(axis * 20 + i) * 2 + 1] = a < b ? b : a;
        }
}

size_t car_frame_map_bytes() { return sizeof(FrameMap); }
size_t car_frame_aux_bytes() { return sizeof(FrameAux); }

cudaError_t car_raster_init() {
    return cudaFuncSetAttribute(car_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RasterSmem));
}

cudaError_t launch_car_stack_shift(const CarDev& p, uint8_t* obs, cudaStream_t s) {
    if (p.ring_mode || p.rot_n > 0 || p.c < 2) return cudaSuccess;
    const int n_units = p.n * p.players * (p.c - 1);
    car_stack_shift_kernel<<<min((n_units + 1) / 2, 2 * 148), SHIFT_THREADS, 0, s>>>(p, obs);
    return cudaGetLastError();
}

// stack mode: the frame ring moves on by one slot (after the render passes AND the stack shift of the step, which read ring_pos)
cudaError_t launch_car_ring_advance(const CarDev& p, cudaStream_t s) {
    if (p.ring_mode || p.rot_n > 0) return cudaSuccess;         // ring mode: the phase is advanced on the host, per step
    car_ring_advance_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_car_render(const CarDev& p, int only_done, int which, int advance, uint8_t* obs, uint8_t* term_obs, cudaStream_t s) {
    // few frames (list passes, small batches): the latency of one thread's chain counts, not the instruction total --
    // one kernel for both parts of the setup, a thread per (polygon, row)
    if (only_done || which == 2 || p.n * p.players <= 4096)
        car_frame_aux_rows_kernel<<<(p.n * p.players + 1) / 2, 2 * AUX_ROW_THREADS, 0, s>>>(p, only_done, which);
    else {
        car_frame_setup_kernel<<<(p.n * p.players + 127) / 128, 128, 0, s>>>(p, only_done, which);
        car_frame_aux_kernel<<<(p.n * p.players * 16 + 127) / 128, 128, 0, s>>>(p, only_done, which);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int ctas = (only_done || which == 2) ? min(p.n * p.players, LIST_PASS_CTAS) : p.n * p.players;
    car_render_kernel<<<ctas, RASTER_THREADS, sizeof(RasterSmem), s>>>(p, only_done, which, obs, term_obs);
    e = cudaGetLastError();
    // ring mode: the phase is advanced on the host, per step; an auto-reset pass sets ring_pos itself
    if (e != cudaSuccess || !advance || p.ring_mode || p.rot_n > 0 || only_done) return e;
    car_ring_advance_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace crl
