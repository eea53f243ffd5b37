// vrt_wave.cuh -- the marcher for INCOHERENT ray batches: one persistent cooperative kernel that re-orders the WORK, round by round.
//
// Randomly directed rays (BASELINE config 4) defeat the single-launch marcher twice over: its warps run nearly empty (rays end
// after 2..4095 steps, a warp that refills only when all of its lanes are idle lives as long as its longest ray) and every cell
// change is a scattered fetch from a 2 GB volume (75 B of DRAM traffic per ray-step, L2 hit rate 20 %).  Here the volume is cut
// into bricks of 2^k voxels per axis and the trace runs in rounds inside ONE kernel launch (grid-wide barriers between the phases
// of a round; nothing returns to the host, the number of rounds is whatever the batch needs):
//     1. bucket pass   counting sort of the rays still alive by the brick they are in: a histogram that the previous round's march
//                      already filled, a grid-wide exclusive scan, and a scatter of the ray ids (the sort need not be stable:
//                      results are written by ray id)
//     2. march         persistent warps pull rays from the brick-sorted list, `refill` lanes at a time, and march every ray until it
//                      has left its brick (plus a margin that keeps rays from bouncing between two bricks) or has finished.  At any
//                      moment the ~100 000 rays in flight on the whole GPU belong to a handful of neighbouring bricks, so a brick is
//                      fetched from DRAM once per round and then served by L2.  A ray that leaves its box writes its exact internal
//                      state (fixed-point position, scaled float direction, step counter, brightness) back and is counted into the
//                      next round's histogram.
// When few rays are left, a last round marches the rest without a box.  Every step is computed exactly like in march3_kernel
// (same operation order: cu:317-374), so results are bit-identical to the single launch.
//
// This replaces round 1's region mode (a cub radix sort and a kernel launch per round, 14 launches + 13 sorts per trace) at the same
// marching rate; what was tried on the way and measured slower is in DESIGN.md section 6 (a CTA per brick with CTA-wide compaction:
// latency-bound; warp work items with CTA-shared batches for L1 affinity: no gain, the front of bricks in flight outgrows L2;
// prefetch.global.L1 one step ahead: 95 -> 56 G ray-steps/s; bricks in Morton instead of row-major order: 91 -> 80; one 256-bit load
// (LDG.E.ENL2.256) for the z-adjacent corner pair of a row where it is naturally aligned: 91 -> 67).
//
// The kernel can be GATED by a device flag written by coherence_probe_kernel, so that vrt_trace_device -- which cannot look at
// device buffers without a synchronisation -- enqueues probe + single-launch marcher + this kernel and exactly one of the two
// marchers does the work.
#pragma once

#include <cooperative_groups.h>

#include "vrt_march.cuh"

namespace vrt {

namespace cg = cooperative_groups;

#ifndef VRT_WAVE_MINCTAS
#define VRT_WAVE_MINCTAS 3     // 3 x 256 threads per SM -> up to 85 registers: no spills (4 CTAs / 64 registers spills the corner cache)
#endif
#ifndef VRT_WAVE_MINCTAS_ALLCLEAR
#define VRT_WAVE_MINCTAS_ALLCLEAR 4   // the all-clear variant keeps 24 instead of 32 cache registers
#endif

constexpr uint32_t kWaveDone = 0xFFFFFFFFu;
constexpr int      kWaveThreads = 256;           // threads per CTA

// control block (device): written by the scan phase, read by everybody after the grid barrier
enum { kCtlAlive = 0, kCtlCursor = 2, kCtlTail = 3, kCtlPrevCount = 4, kCtlRounds = 5, kCtlWords = 8 };

struct WaveParams
{
    MarchParams m;               // volume, limits, invscale, inputs, outputs, n, cap flag; mode_flag/mode_want gate the launch
    uint32_t  *st_pos;           // [n][3] suspended position
    float     *st_dir;           // [n][3] suspended internal direction (already scaled by 2^16 / 2^8)
    uint32_t  *st_it;            // [n]    remaining-iterations counter (the reference's raydata_t::_iterations)
    uint32_t  *st_light;         // [n]    brightness (live translucency only)
    uint32_t  *key_of_ray;       // [n]    brick of the ray after the last round, kWaveDone when finished
    uint32_t  *order[2];         // [n]    ray ids of the rays alive, grouped by brick (ping-pong)
    uint32_t  *hist[2];          // [K]    rays per brick for this / the next round
    uint32_t  *bin_off;          // [K+1]  scatter cursor of every brick (starts at the brick's offset into order[]); [K]: the tail round's single cursor
    uint32_t  *partial;          // [gridDim] per-CTA ray totals of the scan
    uint32_t  *ctl;              // [kCtlWords]
    int        log2_brick;       // brick edge = 2^log2_brick voxels
    uint32_t   margin;           // voxels a ray may travel beyond its brick before it is suspended
    uint32_t   nby, nbz, K;      // bricks along axes 1, 2; number of bricks
    uint32_t   tail_rays;        // when at most this many rays are alive the rest is marched without a box
    uint32_t   refill;           // a warp takes new rays from the list when at least this many lanes are idle
    int        steps_per_check;  // marching steps between two refill polls
    uint32_t   max_rounds;       // safety net (a round always advances every ray by at least one step)
};

__device__ __forceinline__ uint32_t wave_key(uint32_t px, uint32_t py, uint32_t pz, const WaveParams &p)
{
    // rays outside the volume get the key of the nearest brick: the march retires them on its first bounds test
    const uint32_t ix = min(px >> 16, p.m.limx) >> p.log2_brick, iy = min(py >> 16, p.m.limy) >> p.log2_brick, iz = min(pz >> 16, p.m.limz) >> p.log2_brick;
    return (ix * p.nby + iy) * p.nbz + iz;
}

// Device-side coherence probe (the host-side twin is batch_is_incoherent() in vrt_api.cu): are neighbouring rays of the batch
// neighbours in space?  4096 scattered pairs (i, i+1); a pair is incoherent when its start positions are more than 4 voxels
// apart or its directions differ by more than ~25 degrees.  *flag = 1 when most pairs are incoherent, else 0.  One CTA.
__global__ void coherence_probe_kernel(const uint32_t *pos, const void *dir, int dir_i16, unsigned long long n, int dim, uint32_t *flag)
{
    __shared__ unsigned s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const unsigned long long samples = n < 2 ? 0 : (n - 1 < 4096 ? n - 1 : 4096);
    unsigned bad = 0;
    for (unsigned long long j = threadIdx.x; j < samples; j += blockDim.x)
    {
        const unsigned long long i = ((j * 0x9E3779B97F4A7C15ull) >> 11) % (n - 1);
        bool far = false;
        float dot = 0, na = 0, nb = 0;
        for (int d = 0; d < dim; ++d)
        {
            const int32_t dp = (int32_t)(pos[(i + 1) * dim + d] - pos[i * dim + d]);
            if (dp > (4 << 16) || dp < -(4 << 16)) far = true;
            const float a = dir_i16 ? (float)((const short *)dir)[i * dim + d] : ((const float *)dir)[i * dim + d];
            const float b = dir_i16 ? (float)((const short *)dir)[(i + 1) * dim + d] : ((const float *)dir)[(i + 1) * dim + d];
            dot += a * b; na += a * a; nb += b * b;
        }
        if (far || !(dot * dot >= 0.81f * na * nb && dot >= 0)) ++bad;
    }
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    if (threadIdx.x == 0) *flag = (samples > 0 && s_bad * 2 > samples) ? 1u : 0u;
}

// block-wide exclusive scan of one value per thread; returns the block total in `total`
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t &total, uint32_t *s_warp /* [kWaveThreads / 32] */)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += x;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kWaveThreads / 32; ++w)
    {
        const uint32_t c = s_warp[w];
        if ((unsigned)w < warp) before += c;
        sum += c;
    }
    __syncthreads();
    total = sum;
    return before + inc - v;
}

// ALLCLEAR: the scene has no voxel that could make a sample opaque (counted at scene creation, VRT_INFO_ALL_CLEAR), so the stop test
// cu:343 cannot fire and channel 3 is neither kept nor interpolated -- the cell cache is march3_kernel's CornersZ (24 instead of 32
// registers, 24 instead of 30 instructions per sample; same bits, see trilerp_z) and the kernel fits 4 resident CTAs per SM.
template <bool ALLCLEAR> struct WaveBounds { static constexpr int kMinCtas = ALLCLEAR ? VRT_WAVE_MINCTAS_ALLCLEAR : VRT_WAVE_MINCTAS; };

template <typename VoxT, bool DIR_I16, bool LIVE, bool ALLCLEAR>
__global__ void __launch_bounds__(kWaveThreads, (WaveBounds<ALLCLEAR>::kMinCtas)) march3_wave_kernel(const WaveParams p)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const MarchParams &m = p.m;
    if (m.mode_flag != nullptr && *m.mode_flag != m.mode_want) return;       // gated launch: the probe chose the other marcher (uniform over the grid)
    cg::grid_group grid = cg::this_grid();
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + tid, gsize = (unsigned long long)gridDim.x * blockDim.x;

    __shared__ uint32_t s_scan[kWaveThreads / 32];

    // ---- phase 0: every ray into the state arrays, first histogram ---------------------------------------------------------
    for (unsigned long long i = gtid; i < (unsigned long long)p.K; i += gsize) { p.hist[0][i] = 0; p.hist[1][i] = 0; }
    grid.sync();
    for (unsigned long long i = gtid; i < m.n; i += gsize)
    {
        uint32_t px, py, pz; float dx, dy, dz;
        load_ray<DIR_I16>(m, i, px, py, pz, dx, dy, dz);
        p.st_pos[i * 3] = px; p.st_pos[i * 3 + 1] = py; p.st_pos[i * 3 + 2] = pz;
        p.st_dir[i * 3] = dx; p.st_dir[i * 3 + 1] = dy; p.st_dir[i * 3 + 2] = dz;
        p.st_it[i] = m.iterations - 1u;                                                          // cu:333
        if (LIVE) p.st_light[i] = 0xFFFFFFFFu;                                                   // cu:332
        const uint32_t key = wave_key(px, py, pz, p);
        VRT_CHK(key < p.K);
        p.key_of_ray[i] = key;
        atomicAdd(&p.hist[0][key], 1u);
        p.order[1][i] = (uint32_t)i;                                                             // "previous list" of round 0: everybody
    }
    if (gtid == 0) { p.ctl[kCtlPrevCount] = (uint32_t)m.n; p.ctl[kCtlRounds] = 0; }
    grid.sync();

    int cur = 0;            // hist[cur] counts this round's rays; order[cur] is built from order[cur ^ 1]
    for (uint32_t round = 0; round < p.max_rounds; ++round)
    {
        // ---- bucket pass 1a: per-CTA ray totals over the CTA's slice of bricks -------------------------------------------------
        const uint32_t slice = ((p.K + gridDim.x - 1) / gridDim.x + kWaveThreads - 1) / kWaveThreads * kWaveThreads;
        const uint32_t b_lo = min(p.K, blockIdx.x * slice), b_hi = min(p.K, b_lo + slice);
        {
            uint32_t a = 0, total;
            for (uint32_t k = b_lo + tid; k < b_hi; k += kWaveThreads) a += p.hist[cur][k];
            block_scan(a, total, s_scan);
            if (tid == 0) p.partial[blockIdx.x] = total;
        }
        grid.sync();
        // ---- bucket pass 1b: every brick's first slot in order[] (its scatter cursor), control block -----------------------------
        {
            uint32_t a = 0, before, all = 0;
            for (uint32_t c = tid; c < gridDim.x; c += kWaveThreads) { const uint32_t v = p.partial[c]; if (c < blockIdx.x) a += v; all += v; }
            block_scan(a, before, s_scan);                             // rays in the slices before this CTA's
            if (blockIdx.x == 0)
            {
                uint32_t alive;
                block_scan(all, alive, s_scan);
                if (tid == 0)
                {
                    p.bin_off[p.K] = 0;                                // the tail round's single cursor
                    p.ctl[kCtlAlive] = alive; p.ctl[kCtlCursor] = 0;
                    p.ctl[kCtlTail] = alive <= p.tail_rays ? 1u : 0u;
                    p.ctl[kCtlRounds] = round + 1;
                }
            }
            uint32_t off = before;
            for (uint32_t k0 = b_lo; k0 < b_hi; k0 += kWaveThreads)
            {
                const uint32_t k = k0 + tid;
                const uint32_t c = k < b_hi ? p.hist[cur][k] : 0u;
                uint32_t tile_total;
                const uint32_t ex = block_scan(c, tile_total, s_scan);
                if (k < b_hi)
                {
                    p.bin_off[k] = off + ex;
                    p.hist[cur ^ 1][k] = 0;                            // the next round's histogram starts empty
                }
                off += tile_total;
            }
        }
        grid.sync();
        const uint32_t alive = p.ctl[kCtlAlive], prev_count = p.ctl[kCtlPrevCount];
        const bool tail = p.ctl[kCtlTail] != 0u;
        if (alive == 0) break;                                                                   // uniform over the grid
        // ---- bucket pass 2: scatter the ids of the rays still alive into their bricks' slots -------------------------------------
        // (tail round: one list, in whatever order -- slot = a single global cursor, kept in bin_off[K])
        for (unsigned long long i = gtid; i < prev_count; i += gsize)
        {
            const uint32_t ray = p.order[cur ^ 1][i];
            const uint32_t k = p.key_of_ray[ray];
            VRT_CHK(ray < m.n && (k == kWaveDone || k < p.K));
            if (k != kWaveDone)
            {
                const uint32_t slot = atomicAdd(&p.bin_off[tail ? p.K : k], 1u);
                VRT_CHK(slot < alive);
                p.order[cur][slot] = ray;
            }
        }
        grid.sync();
        // ---- march: persistent warps pull rays from the brick-sorted list ------------------------------------------------------
        // Consecutive warps -- on all SMs -- work on the same few bricks at the same time, so a brick is fetched from DRAM once per
        // round and then served by L2 (and L1).  A ray marches inside its brick's box (the box is per lane); idle lanes are refilled
        // from the list as soon as `refill` of them are free, so warps stay full although rays leave their bricks after very
        // different numbers of steps.
        {
            const float invx = m.invx, invy = m.invy, invz = m.invz;
            uint32_t lo_x = 0, lo_y = 0, lo_z = 0, sp_x = 0, sp_y = 0, sp_z = 0;       // this lane's box: [lo, lo + span) in 16.16, clipped to the volume
            uint32_t px = 0, py = 0, pz = 0, it = 0, brightness = 0xFFFFFFFFu, cached_tr = 0, moved = 0xFFFFFFFFu, ray = 0;
            float dx = 0, dy = 0, dz = 0;
            bool have = false, exhausted = false;                                    // exhausted: warp-uniform, the list has been handed out
            CornersP q;             // cell cache (generic)
            CornersZ cz;            // cell cache without channel 3 (ALLCLEAR)
            for (;;)
            {
                if (!exhausted)
                {
                    const unsigned idle = __ballot_sync(FULL, !have);
                    const uint32_t nidle = (uint32_t)__popc(idle);
                    if (nidle >= p.refill)
                    {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&p.ctl[kCtlCursor], nidle);
                        base = __shfl_sync(FULL, base, 0);
                        if (!have)
                        {
                            const uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                            if (idx < alive)
                            {
                                ray = p.order[cur][idx];
                                VRT_CHK(ray < m.n);
                                px = p.st_pos[(size_t)ray * 3]; py = p.st_pos[(size_t)ray * 3 + 1]; pz = p.st_pos[(size_t)ray * 3 + 2];
                                dx = p.st_dir[(size_t)ray * 3]; dy = p.st_dir[(size_t)ray * 3 + 1]; dz = p.st_dir[(size_t)ray * 3 + 2];
                                it = p.st_it[ray];
                                if (LIVE) brightness = p.st_light[ray];
                                moved = 0xFFFFFFFFu;
                                if (tail) { lo_x = lo_y = lo_z = 0; sp_x = m.limx16; sp_y = m.limy16; sp_z = m.limz16; }
                                else
                                {
                                    // the brick the ray is in (the same arithmetic that produced its key), widened by the margin
                                    const uint32_t e = 1u << p.log2_brick;
                                    const uint32_t ax = (min(px >> 16, m.limx) >> p.log2_brick) * e, ay = (min(py >> 16, m.limy) >> p.log2_brick) * e,
                                                   az = (min(pz >> 16, m.limz) >> p.log2_brick) * e;
                                    const uint32_t l_x = ax > p.margin ? ax - p.margin : 0u, l_y = ay > p.margin ? ay - p.margin : 0u, l_z = az > p.margin ? az - p.margin : 0u;
                                    const uint32_t h_x = min(ax + e + p.margin, m.limx), h_y = min(ay + e + p.margin, m.limy), h_z = min(az + e + p.margin, m.limz);
                                    lo_x = l_x << 16; lo_y = l_y << 16; lo_z = l_z << 16;
                                    sp_x = (h_x - l_x) << 16; sp_y = (h_y - l_y) << 16; sp_z = (h_z - l_z) << 16;
                                }
                                have = true;
                            }
                        }
                        if (base + nidle >= alive) exhausted = true;
                    }
                }
                if (!__any_sync(FULL, have)) { if (exhausted) break; else continue; }
                if (have)
                {
                    const uint32_t it_stop = it - min(it, (uint32_t)p.steps_per_check);
                    bool done = false, suspend = false;
                    uint32_t it_final = 0;
                    while (it != it_stop)
                    {
                        if (!(((px - lo_x) < sp_x) & ((py - lo_y) < sp_y) & ((pz - lo_z) < sp_z)))
                        {
                            // outside this round's box: either outside the volume (the reference's loop condition fails, cu:335) or only
                            // outside the brick (suspend; the next round continues with exactly this state)
                            if (!((px < m.limx16) & (py < m.limy16) & (pz < m.limz16))) { done = true; it_final = it; }
                            else suspend = true;
                            break;
                        }
                        --it;
                        if (moved >= 0x10000u)
                        {
                            const uint32_t cell = ((px >> 16) * m.by + (py >> 16)) * m.bz + (pz >> 16);       // cu:113
                            VRT_CHK(cell < m.nvox);
                            if (LIVE) cached_tr = 0xFFFFFFFFu - ldg_nc_u32(m.translucency + cell);
                            if (ALLCLEAR) load_corners_z<VoxT>(cz, m, cell);
                            else          load_corners<VoxT>(q, m, cell);
                        }
                        if (LIVE)                                                                            // cu:337-341
                        {
                            const uint32_t absorb = cached_tr;
                            brightness -= min(brightness, absorb);
                            if (brightness < m.min_brightness) { done = true; it_final = it + 1u; break; }
                        }
                        unsigned long long gxy, gzw;
                        float gz, gw, sx, sy;
                        if (ALLCLEAR) trilerp_z(cz, px, py, pz, gxy, gz, scale48_const());                   // cu:342; cu:343 cannot fire
                        else
                        {
                            trilerp_packed(q, px, py, pz, gxy, gzw, scale48_const());                        // cu:342
                            unpack2(gzw, gz, gw);
                            if (gw > 0.0f) { done = true; it_final = it + 1u; break; }                       // cu:343
                        }
                        const unsigned long long dxy = fma2(pack2(invx, invy), gxy, pack2(dx, dy));          // cu:344-345
                        dz = __fmaf_rn(invz, gz, dz);
                        unpack2(dxy, dx, dy);
                        const float dot = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                        // cu:346-347 with the two exact shortcuts of march3_kernel's fast loop (the short division sequence and the
                        // add-a-constant rounding, each proven equal to div.rn.f32 / cvt.rni over the whole |dir|^2 range the host
                        // puts into dot_lo / dot_span: vrt_selftest_division); outside that range the original instructions run
                        const unsigned long long sdxy = mul2(pack2(invx, invy), dxy);
                        const float sdz = __fmul_rn(invz, dz);
                        float ilen = div_fast(dot);
                        unpack2(mul2(sdxy, pack2(ilen, ilen)), sx, sy);
                        uint32_t ax = rni_small(sx), ay = rni_small(sy), az = rni_small(__fmul_rn(sdz, ilen));
                        if (!div_is_fast_in(dot, m.dot_lo, m.dot_span))
                        {
                            ilen = __fdiv_rn(0x42000000p0f, dot);                                            // cu:346
                            unpack2(mul2(sdxy, pack2(ilen, ilen)), sx, sy);                                  // cu:347
                            ax = (uint32_t)__float2int_rn(sx); ay = (uint32_t)__float2int_rn(sy); az = (uint32_t)__float2int_rn(__fmul_rn(sdz, ilen));
                        }
                        const uint32_t nx = px + ax, ny = py + ay, nz = pz + az;
                        moved = (px ^ nx) | (py ^ ny) | (pz ^ nz);
                        px = nx; py = ny; pz = nz;
                    }
                    if (!done && !suspend && it == 0u) { done = true; it_final = 0u; }                       // cap (cu:335,350)
                    if (done)
                    {
                        store_ray<DIR_I16, LIVE, false>(m, ray, px, py, pz, dx, dy, dz, it_final, brightness);
                        p.key_of_ray[ray] = kWaveDone;
                        have = false;
                    }
                    else if (suspend)
                    {
                        p.st_pos[(size_t)ray * 3] = px; p.st_pos[(size_t)ray * 3 + 1] = py; p.st_pos[(size_t)ray * 3 + 2] = pz;
                        p.st_dir[(size_t)ray * 3] = dx; p.st_dir[(size_t)ray * 3 + 1] = dy; p.st_dir[(size_t)ray * 3 + 2] = dz;
                        p.st_it[ray] = it;
                        if (LIVE) p.st_light[ray] = brightness;
                        const uint32_t key = wave_key(px, py, pz, p);
                        VRT_CHK(key < p.K);
                        p.key_of_ray[ray] = key;
                        atomicAdd(&p.hist[cur ^ 1][key], 1u);
                        have = false;
                    }
                }
            }
        }
        if (gtid == 0) p.ctl[kCtlPrevCount] = alive;
        grid.sync();
        cur ^= 1;
    }
}

} // namespace vrt
