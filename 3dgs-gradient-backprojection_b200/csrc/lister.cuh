// The "lister" warp of the two tcgen05 back-projection kernels (backproject_tc.cu, backproject_lr.cu).
//
// It owns the work queue (tiles are pulled from a global counter) and turns a tile's depth-ordered list into batches of
// <= 128 packed Gaussian indices in a small shared-memory ring, running a few batches ahead of the ALU warps -- across
// tile boundaries, so a tile switch costs the ALU warps nothing.  Two list kinds (TileCtx):
//   * per-tile lists (gsplat's flatten_ids / isect_offsets): the batch is a copy of 128 consecutive ids;
//   * per-supertile lists (GWBP_PREPARE_SUPERTILE): the supertile's entries (packed index, mask of its 8 x 4 tiles)
//     are streamed 64 per step and the ones whose mask holds this tile's bit are compacted in list order -- same
//     Gaussians, same order as the per-tile list would hold, at ~1/2.4 of the entries to emit and sort up front.
// Every tile ends with a batch flagged `last` (possibly empty).  When the ALU warps find every pixel of a tile finished
// they set `abort_unit`; the lister stops scanning that tile and the ALU warps skip the tile's remaining batches.  After
// its last tile a lister publishes a batch with unit = -1.
//
// TWO listers, each with its own ring, take alternate tiles: while the ALU warps work on a tile from ring r, the other
// lister prepares the next tile's first batches, so neither the tile switch (work-queue atomic -> list bounds -> entries:
// three dependent round trips, then ~1 700 entries scanned for the first batch) nor the lag between the ALU warps
// finishing a tile early and its lister noticing is ever on the ALU warps' path.  [One lister: bp_tc 1.03 -> 1.32 ms.]
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace gwbp {
namespace lst {

constexpr int NL = 2;     // id batches in flight between one lister and the ALU warps
constexpr int NLISTERS = 2;
constexpr int LB = 128;   // ids per batch = Gaussians per MMA batch

struct alignas(8) Ring {
    int ids[NL][LB];
    int n[NL], unit[NL], last[NL];
    volatile int abort_unit;
};

// Tile visiting order.  Work units are handed out in bands of `kband` tile rows, column-major inside a band, so that the
// ~148 tiles in flight form a compact block: the (typically 2x2..3x3) tiles that touch one Gaussian are processed close
// together in time and their row reductions merge in the 126 MB L2 instead of each costing a DRAM read-modify-write of the
// accumulator row.
__device__ __forceinline__ int unit_to_tile(int unit, int tw, int th, int kband) {
    const int per_band = kband * tw;
    const int band = unit / per_band, r = unit - band * per_band;
    const int hb = min(kband, th - band * kband);
    const int tx = r / hb, ty = band * kband + (r - tx * hb);
    return ty * tw + tx;
}

// full[i] = bars + 8 * i (count 1), free[i] = bars + 8 * (NL + i) (count = number of consumer warps): shared-memory
// addresses of 2 * NL mbarriers initialised by the caller.
__device__ __forceinline__ void run_lister(const TileCtx &t, int *unit_counter, int nunits, int band, Ring *R,
                                           uint32_t bars) {
    using namespace tc;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int ql = 0, slot = 0;
    auto begin_batch = [&]() {
        slot = ql % NL;
        if (ql >= NL) mbar_wait(bars + 8 * (NL + slot), ((ql / NL) - 1) & 1);
    };
    auto publish = [&](int unit, int n, int last) {
        __syncwarp();
        if (lane == 0) {
            R->n[slot] = n;
            R->unit[slot] = unit;
            R->last[slot] = last;
            mbar_arrive(bars + 8 * slot);
        }
        ++ql;
    };
    while (true) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(unit_counter, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= nunits) break;
        const int tile = unit_to_tile(unit, t.tw, t.th, band);
        const int ty = tile / t.tw, tx = tile - ty * t.tw;
        begin_batch();
        if (t.sents == nullptr) {
            // ---- per-tile list: batches are plain copies
            const int s = t.offsets[tile], e = t.offsets[tile + 1];
            int b = s;
            while (true) {
                const int n = min(LB, e - b);
#pragma unroll
                for (int j = 0; j < LB / 32; ++j)
                    if (32 * j + lane < n) R->ids[slot][32 * j + lane] = t.flatten[b + 32 * j + lane];
                const bool last = b + LB >= e || __shfl_sync(0xffffffffu, (int)(R->abort_unit == unit), 0) != 0;
                publish(unit, max(n, 0), last ? 1 : 0);
                if (last) break;
                b += LB;
                begin_batch();
            }
            continue;
        }
        // ---- supertile list: keep the entries whose tile mask holds this tile's bit
        const int st = (ty / kSuperH) * t.nsx + tx / kSuperW;
        const int kbit = (ty % kSuperH) * kSuperW + tx % kSuperW;
        const long long s = t.offsets[st], e = t.offsets[st + 1];
        const uint2 *ents = t.sents;
        auto ld = [&](long long cc) -> uint4 {  // entries cc + 2 * lane, + 1 (16-byte aligned: cc is even)
            const long long idx = cc + 2 * lane;
            return idx < e ? __ldg(reinterpret_cast<const uint4 *>(ents + idx)) : make_uint4(0u, 0u, 0u, 0u);
        };
        long long c = s & ~1ll;
        uint4 pre0 = ld(c), pre1 = ld(c + 64), pre2 = ld(c + 128), pre3 = ld(c + 192);
        int found = 0;
        for (; c < e; c += 64) {
            const uint4 v = pre0;
            pre0 = pre1; pre1 = pre2; pre2 = pre3;
            pre3 = ld(c + 256);
            if ((c & 255) == 0 && __shfl_sync(0xffffffffu, (int)(R->abort_unit == unit), 0) != 0) {
                found = 0;  // the ALU warps are done with this tile
                break;
            }
            const long long i0 = c + 2 * lane;
            const bool h0 = i0 >= s && i0 < e && ((v.y >> kbit) & 1u);
            const bool h1 = i0 + 1 < e && ((v.w >> kbit) & 1u);  // i0 + 1 >= s always holds (c >= s - 1)
            const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
            if ((m0 | m1) == 0u) continue;
            const int r0 = found + __popc(m0 & lt) + __popc(m1 & lt), r1 = r0 + (h0 ? 1 : 0);
            const int total = found + __popc(m0) + __popc(m1);
            if (h0 && r0 < LB) R->ids[slot][r0] = (int)v.x;
            if (h1 && r1 < LB) R->ids[slot][r1] = (int)v.z;
            if (total >= LB) {
                publish(unit, LB, 0);
                const bool aborted = __shfl_sync(0xffffffffu, (int)(R->abort_unit == unit), 0) != 0;
                begin_batch();
                if (aborted) {
                    found = 0;
                    break;
                }
                if (h0 && r0 >= LB) R->ids[slot][r0 - LB] = (int)v.x;
                if (h1 && r1 >= LB) R->ids[slot][r1 - LB] = (int)v.z;
                found = total - LB;
            } else {
                found = total;
            }
        }
        publish(unit, found, 1);
    }
    begin_batch();
    publish(-1, 0, 1);
}

}  // namespace lst
}  // namespace gwbp
