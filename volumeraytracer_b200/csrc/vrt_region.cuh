// vrt_region.cuh -- opt-in marching mode for INCOHERENT ray batches: L2-resident regions.
//
// With randomly directed rays every cell change is a DRAM access (config 4: 75 B of DRAM traffic per ray-step, L2 hit
// rate 20 %): the gathers of the rays in flight are spread over the whole volume.  This mode re-orders the work instead of
// the data: the volume is cut into cubic regions of 2^k voxels (k = VRT_OPT_REGION_LOG2; 128^3 float4 voxels = 34 MB, a
// fraction of the 126 MB L2), rays are sorted by the region they are in (cub radix sort on a 16-bit key), and ONE persistent
// launch marches every ray until it leaves its region (plus a margin that stops rays from bouncing between two regions) or
// terminates.  Consecutive warps work on the same region, so its voxels are fetched from DRAM once and then served by L2.
// Suspended rays keep their exact internal state (fixed-point position, scaled float direction, step counter, brightness)
// in global memory, so every step is computed exactly as in the single-launch marcher: results are bit-identical.
// A fixed number of rounds is enqueued (rounds with nothing left to do exit at once) followed by one unrestricted round, so
// the whole trace stays asynchronous on the caller's stream.
#pragma once

#include "vrt_march.cuh"

namespace vrt {

constexpr uint32_t kRegionDone = 0xFFFFu;

struct RegionParams
{
    MarchParams m;               // volume, limits, invscale, outputs (epos/edir/eit/light), n, counter, refill, steps_per_poll
    uint32_t  *st_pos;           // [n][3] suspended position
    float     *st_dir;           // [n][3] suspended internal direction (already scaled by 2^16 / 2^8)
    uint32_t  *st_it;            // [n]    remaining-iterations counter (the reference's raydata_t::_iterations)
    uint32_t  *st_light;         // [n]    brightness (live translucency only)
    const uint32_t *order;       // [n]    ray ids sorted by region key
    uint16_t  *keys;             // [n]    in: region key of order[k]; out: its new key (kRegionDone when the ray has finished)
    int        log2_edge;        // region edge = 2^log2_edge voxels; < 0: unrestricted round
    uint32_t   margin;           // voxels a ray may travel beyond its region before it is suspended
    uint32_t   ry, rz;           // regions along axes 1, 2 (key = (rx*ry + ry_)*rz + rz_)
};

__device__ __forceinline__ uint16_t region_key(uint32_t px, uint32_t py, uint32_t pz, int log2_edge, uint32_t ry, uint32_t rz,
                                               uint32_t limx, uint32_t limy, uint32_t limz)
{
    // rays outside the volume get the key of the nearest region: the marcher then retires them on its first bounds test
    const uint32_t ix = min(px >> 16, limx), iy = min(py >> 16, limy), iz = min(pz >> 16, limz);
    return (uint16_t)(((ix >> log2_edge) * ry + (iy >> log2_edge)) * rz + (iz >> log2_edge));
}

template <bool DIR_I16>
__global__ void region_init_kernel(const RegionParams p, uint32_t *order_init)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m.n) return;
    uint32_t px, py, pz; float dx, dy, dz;
    load_ray<DIR_I16>(p.m, i, px, py, pz, dx, dy, dz);
    p.st_pos[i * 3] = px; p.st_pos[i * 3 + 1] = py; p.st_pos[i * 3 + 2] = pz;
    p.st_dir[i * 3] = dx; p.st_dir[i * 3 + 1] = dy; p.st_dir[i * 3 + 2] = dz;
    p.st_it[i] = p.m.iterations - 1u;                                                          // cu:333
    p.st_light[i] = 0xFFFFFFFFu;                                                               // cu:332
    p.keys[i] = region_key(px, py, pz, p.log2_edge, p.ry, p.rz, p.m.limx, p.m.limy, p.m.limz);
    order_init[i] = (uint32_t)i;
}

template <typename VoxT, bool DIR_I16, bool LIVE>
__global__ void __launch_bounds__(128, 5) march3_region_kernel(const RegionParams p)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const MarchParams &m = p.m;

    uint32_t px = 0, py = 0, pz = 0, it = 0, brightness = 0xFFFFFFFFu, cached_tr = 0, moved = 0xFFFFFFFFu;
    uint32_t lo_x = 0, lo_y = 0, lo_z = 0, sp_x = 0, sp_y = 0, sp_z = 0;   // this ray's box: [lo, lo + span) in 16.16, clipped to the volume
    float dx = 0, dy = 0, dz = 0;
    unsigned long long slot = 0;
    uint32_t ray = 0;
    bool have = false, exhausted = false;
    CornersP q;
    const float invx = m.invx, invy = m.invy, invz = m.invz;

    for (;;)
    {
        if (!exhausted)
        {
            const unsigned idle = __ballot_sync(FULL, !have);
            const int nidle = __popc(idle);
            if (nidle >= m.refill)
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(m.counter, (unsigned long long)nidle);
                base = __shfl_sync(FULL, base, 0);
                bool saw_done = false;
                if (!have)
                {
                    const unsigned long long k = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                    if (k < m.n)
                    {
                        const uint32_t key = p.keys[k];
                        if (key == kRegionDone) saw_done = true;        // keys are sorted: everything from here on has finished
                        else
                        {
                            slot = k; ray = p.order[k];
                            px = p.st_pos[(size_t)ray * 3]; py = p.st_pos[(size_t)ray * 3 + 1]; pz = p.st_pos[(size_t)ray * 3 + 2];
                            dx = p.st_dir[(size_t)ray * 3]; dy = p.st_dir[(size_t)ray * 3 + 1]; dz = p.st_dir[(size_t)ray * 3 + 2];
                            it = p.st_it[ray];
                            if (LIVE) brightness = p.st_light[ray];
                            moved = 0xFFFFFFFFu;
                            // the box this ray may march in during this round
                            if (p.log2_edge < 0) { lo_x = lo_y = lo_z = 0; sp_x = m.limx16; sp_y = m.limy16; sp_z = m.limz16; }
                            else
                            {
                                const uint32_t e = 1u << p.log2_edge;
                                const uint32_t cx = min(px >> 16, m.limx) >> p.log2_edge, cy = min(py >> 16, m.limy) >> p.log2_edge, cz = min(pz >> 16, m.limz) >> p.log2_edge;
                                const uint32_t ax = cx * e, ay = cy * e, az = cz * e;
                                const uint32_t l_x = ax > p.margin ? ax - p.margin : 0u, l_y = ay > p.margin ? ay - p.margin : 0u, l_z = az > p.margin ? az - p.margin : 0u;
                                const uint32_t h_x = min(ax + e + p.margin, m.limx), h_y = min(ay + e + p.margin, m.limy), h_z = min(az + e + p.margin, m.limz);
                                lo_x = l_x << 16; lo_y = l_y << 16; lo_z = l_z << 16;
                                sp_x = (h_x - l_x) << 16; sp_y = (h_y - l_y) << 16; sp_z = (h_z - l_z) << 16;
                            }
                            have = true;
                        }
                    }
                }
                if (base + (unsigned long long)nidle >= m.n || __any_sync(FULL, saw_done)) exhausted = true;
            }
        }
        if (!__any_sync(FULL, have)) break;
        if (!have) continue;

        const uint32_t it_stop = it - min(it, (uint32_t)m.steps_per_poll);
        bool done = false, suspend = false;
        uint32_t it_final = 0;
        while (it != it_stop)
        {
            if (!(((px - lo_x) < sp_x) & ((py - lo_y) < sp_y) & ((pz - lo_z) < sp_z)))
            {
                // outside this round's box: either outside the volume (the reference's loop condition fails, cu:335) or only
                // outside the region (suspend; the next round continues with exactly this state)
                if (!((px < m.limx16) & (py < m.limy16) & (pz < m.limz16))) { done = true; it_final = it; }
                else suspend = true;
                break;
            }
            --it;
            if (moved >= 0x10000u)
            {
                const uint32_t cell = ((px >> 16) * m.by + (py >> 16)) * m.bz + (pz >> 16);       // cu:113
                if (LIVE) cached_tr = ldg_nc_u32(m.translucency + cell);
                if (m.pair)       load_corners_pair(q, m, cell);     // uniform branches
                else if (m.brick) load_corners_brick<VoxT>(q, m.volume, px >> 16, py >> 16, pz >> 16, m.nby, m.nbz);
                else              load_corners<VoxT>(q, m, cell);
            }
            if (LIVE)                                                                            // cu:337-341
            {
                const uint32_t absorb = 0xFFFFFFFFu - cached_tr;
                brightness -= min(brightness, absorb);
                if (brightness < m.min_brightness) { done = true; it_final = it + 1u; break; }
            }
            unsigned long long gxy, gzw;
            float gz, gw, sx, sy;
            trilerp_packed(q, px, py, pz, gxy, gzw, scale48_const());                            // cu:342
            unpack2(gzw, gz, gw);
            if (gw > 0.0f) { done = true; it_final = it + 1u; break; }                           // cu:343
            unsigned long long dxy = fma2(pack2(invx, invy), gxy, pack2(dx, dy));                // cu:344-345
            dz = __fmaf_rn(invz, gz, dz);
            unpack2(dxy, dx, dy);
            const float dot = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float ilen = __fdiv_rn(0x42000000p0f, dot);                                    // cu:346
            unpack2(mul2(mul2(pack2(invx, invy), dxy), pack2(ilen, ilen)), sx, sy);              // cu:347
            const uint32_t nx = px + (uint32_t)__float2int_rn(sx);
            const uint32_t ny = py + (uint32_t)__float2int_rn(sy);
            const uint32_t nz = pz + (uint32_t)__float2int_rn(__fmul_rn(__fmul_rn(invz, dz), ilen));
            moved = (px ^ nx) | (py ^ ny) | (pz ^ nz);
            px = nx; py = ny; pz = nz;
        }
        if (!done && !suspend && it == 0u) { done = true; it_final = 0u; }                       // cap (cu:335,350)
        if (done)
        {
            store_ray<DIR_I16, LIVE, false>(m, ray, px, py, pz, dx, dy, dz, it_final, brightness);
            p.keys[slot] = (uint16_t)kRegionDone;
            have = false;
        }
        else if (suspend)
        {
            p.st_pos[(size_t)ray * 3] = px; p.st_pos[(size_t)ray * 3 + 1] = py; p.st_pos[(size_t)ray * 3 + 2] = pz;
            p.st_dir[(size_t)ray * 3] = dx; p.st_dir[(size_t)ray * 3 + 1] = dy; p.st_dir[(size_t)ray * 3 + 2] = dz;
            p.st_it[ray] = it;
            if (LIVE) p.st_light[ray] = brightness;
            p.keys[slot] = region_key(px, py, pz, p.log2_edge, p.ry, p.rz, m.limx, m.limy, m.limz);
            have = false;
        }
    }
}

} // namespace vrt
