// K1 for the TWO innermost axes in one read: get_bitinformation(ds) with dim=None loops the dimensions and runs
// the backend once per dimension (xbitinfo/xbitinfo.py:387-399) -- four passes over a 4-d variable.  A column
// walker (xbi_count_cols.cu) already holds, for every element, the element above it (previous row, in
// registers) and -- but for the last word of its 16-byte chunk -- the element to its right.  This kernel counts
// both adjacencies while the data is there: three bit-sliced counters (sum x', sum x' & above, sum x' & right)
// instead of 2 + 2, one transform per element instead of two, and half the HBM traffic of two passes.
//
//   rows      R = product of all extents but the last, `nx` elements each (row_bytes a multiple of 16)
//   y pairs   (row r-1, row r): every row but the first is a `cur` row once; the walker formulation of
//             xbi_count_cols.cu -- b-stream and a&b counted, a = b + first row - last row (row-plane passes),
//             pairs that straddle two blocks of the second-to-last axis subtracted by k_block_edges
//   x pairs   (e, e+1) inside every `cur` row.  The right neighbour of a lane's last word is the first word of the
//             next lane's chunk: read from the warp's slice of the shared-memory ring (every lane copies its own
//             16 bytes with cp.async; lane 31 copies one word more) after a __syncwarp, and transformed once more.
//             With S = sum of x' over the cur rows: A_x = S - (last column), B_x = S - (first column); the one lane
//             per row that holds a column end sums it in a small counter of its own (a strided pass over the two
//             columns cost 13 % of the call), row 0 of the matrix goes through the generic kernel.
//
// Masked modes: the kernel runs the unmasked formulation and only PROBES for masked elements (x*0 on the FMA pipe,
// xbi_mask.cuh); when it finds one it raises a flag and the caller recounts the two axes one by one with the
// masked kernels.  Data without masked elements -- the API's default NaN mask on NaN-free data -- pays one pass.
// 4-byte element types; everything else is counted axis by axis.
#include <atomic>
#include <type_traits>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"
#include "xbi_mask.cuh"
#include "xbi_tma.cuh"

#pragma nv_diag_suppress 177

namespace xbi {

namespace {

constexpr int kThreadsXY = 512;
constexpr int kWarpsXY = kThreadsXY / 32;
constexpr int kRowPitch = 512 + 16;        // a warp's slice of a row + lane 31's extra word (16-byte slot)
constexpr int kTileXY = 4 * kRowPitch;     // one 4-row group of a warp
constexpr int kRingXY = 6;
constexpr int kAheadXY = kRingXY - 1;
constexpr size_t kSmemXY = 128 + (size_t)kWarpsXY * kRingXY * kTileXY;

__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KIND, bool XFORM>
__device__ __forceinline__ uint32_t xf(uint32_t w) {
    if (XFORM && KIND == kF32) return sexp_f32(w);
    return w;
}

// Work item = (row segment, strip of 32 chunks of a row); a warp takes one item at a time (grid-stride), so all
// its lanes walk the same rows.  Lanes past the end of the row idle (their ring cells hold the neutral value, so
// the lane to their left sees "no right neighbour" without a test).
template <int KIND, bool XFORM, int MT>
__global__ void __launch_bounds__(kThreadsXY, 1)
k_pairs_cols_xy(const char* __restrict__ x, long long row_bytes, long long cpr, long long spr, long long nseg,
                long long seg_min, long long seg_extra, uint32_t mask_lo, uint32_t mask_hi,
                unsigned long long y_flat_pairs, unsigned long long x_pairs, Workspace* __restrict__ ws_y,
                Workspace* __restrict__ ws_x, unsigned int* __restrict__ dirty_flag) {
    using L = Ladder<3, 8>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_addr(smem_raw) + 127u) & ~127u;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wbase = smem0 + warp * (kRingXY * kTileXY);
    const uint32_t my0 = wbase + lane * 16;  // + slot * kTileXY + r * kRowPitch
    constexpr uint32_t kNeutral = (XFORM && KIND == kF32) ? 0x3F800000u : 0u;

    const long long n_items = spr * nseg;
    const long long warps_total = (long long)gridDim.x * kWarpsXY;
    // 32-bit loop counters and opaque loop invariants: left alone, ptxas re-derives `groups` from the kernel
    // parameters and compares 64-bit counters in every group (a dozen uniform-pipe instructions per iteration)
    unsigned groups = (unsigned)(seg_min >> 2);
    asm volatile("" : "+r"(groups));

    // block totals (flushes are rare: they add straight into shared memory, no per-lane 64-bit accumulators)
    __shared__ unsigned long long s_sum[5][32];  // sum x' (cur rows), x' & above, x' & right, first column, last column
    for (int i = threadIdx.x; i < 5 * 32; i += kThreadsXY) (&s_sum[0][0])[i] = 0;
    __syncthreads();

    L Bl, ABy, ABx;
    Bl.clear(); ABy.clear(); ABx.clear();
    uint32_t it = 0;
    bool dirty = false;
    // first / last column of the rows (A_x = S - last column, B_x = S - first column): the one lane of a row that
    // holds it feeds one word per row into a small counter of its own (4 rows = one chunk); warps whose strip
    // touches neither end of the row skip it altogether
    using LC = Ladder<3, 5, 2>;
    LC Col;
    Col.clear();
    uint32_t cit = 0, role = 0;  // role of this lane in the current item: 1 first column, 2 last column
    auto flush_col = [&]() {
        if (cit == 0) return;  // (warp-uniform)
        LC t = Col;
        if (role != 1u) t.clear();
        const uint32_t f = t.warp_totals(lane, cit);
        if (f) atomicAdd(&s_sum[3][lane], (unsigned long long)f);
        t = Col;
        if (role != 2u) t.clear();
        const uint32_t l = t.warp_totals(lane, cit);
        if (l) atomicAdd(&s_sum[4][lane], (unsigned long long)l);
        Col.clear();
        cit = 0;
    };
    auto flush = [&](bool keep) {
        // (an idle lane's counters hold zeros: its words were neutral or never added)
        const uint32_t b = Bl.warp_totals(lane, it), y = ABy.warp_totals(lane, it), xx = ABx.warp_totals(lane, it);
        if (keep) {
            if (b) atomicAdd(&s_sum[0][lane], (unsigned long long)b);
            if (y) atomicAdd(&s_sum[1][lane], (unsigned long long)y);
            if (xx) atomicAdd(&s_sum[2][lane], (unsigned long long)xx);
        }
        Bl.clear(); ABy.clear(); ABx.clear();
        it = 0;
    };
    uint32_t slot_w = 0, slot_r = 0;

    for (long long wi = (long long)blockIdx.x * kWarpsXY + warp; wi < n_items; wi += warps_total) {
        const long long seg = wi / spr, strip = wi - seg * spr;
        const long long chunk = strip * 32 + lane;
        const bool live = chunk < cpr;
        const bool has_right_word = lane == 31u && chunk + 1 < cpr;  // lane 31 fetches the word after the strip
        const uint32_t new_role = (live && chunk == 0) ? 1u : (live && chunk == cpr - 1) ? 2u : 0u;
        if (__any_sync(0xFFFFFFFFu, new_role != role)) flush_col();
        role = new_role;
        const bool col_warp = __any_sync(0xFFFFFFFFu, role != 0u);
        const long long row_begin = seg * seg_min + (seg < seg_extra ? seg : seg_extra);
        int tail_rows = (int)(seg_min + (seg < seg_extra ? 1 : 0) - ((long long)groups << 2));  // 0..4, warp-uniform
        unsigned ngroups = groups + (tail_rows > 0 ? 1u : 0u);
        asm volatile("" : "+r"(tail_rows), "+r"(ngroups));

        // cells the copies of this item will never write: neutral value in every slot of the ring
        // (nothing of the previous item is in flight any more)
        if (!live || (lane == 31u && !has_right_word)) {
#pragma unroll 1
            for (int c = 0; c < kRingXY * 4; ++c) {
                const uint32_t cell = my0 + (c >> 2) * kTileXY + (c & 3) * kRowPitch + (live ? 16u : 0u);
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(cell), "r"(kNeutral) : "memory");
            }
        }
        const char* q0 = x + row_begin * row_bytes + (live ? chunk : 0) * 16;
        const char* q = q0 + row_bytes;
        unsigned issued = 0;
        auto issue = [&]() {
            if (issued < ngroups && live) {
                const uint32_t dst = my0 + slot_w * kTileXY;
                if (issued < groups) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        cpa16(dst + r * kRowPitch, q + r * row_bytes);
                        if (has_right_word) cpa4(dst + r * kRowPitch + 16, q + r * row_bytes + 16);
                    }
                    q += 4 * row_bytes;
                } else {
                    // the tail group re-reads an in-bounds row for its absent rows (they are masked out below)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const long long rr = (r < tail_rows) ? r : (long long)tail_rows - 1;
                        cpa16(dst + r * kRowPitch, q + rr * row_bytes);
                        if (has_right_word) cpa4(dst + r * kRowPitch + 16, q + rr * row_bytes + 16);
                    }
                }
            }
            cpa_commit();
            slot_w = (slot_w + 1 == kRingXY) ? 0 : slot_w + 1;
            ++issued;
        };
#pragma unroll 1
        for (int k = 0; k < kAheadXY; ++k) issue();
        uint32_t prev[4];
        {
            uint4 f = make_uint4(kNeutral, kNeutral, kNeutral, kNeutral);
            if (live) f = __ldg(reinterpret_cast<const uint4*>(q0));
            uint32_t raw[4] = {f.x, f.y, f.z, f.w};
            if (MT != kMaskNone) dirty |= may_contain_masked<KIND, MT>(raw, mask_lo, mask_hi);
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[i] = xf<KIND, XFORM>(raw[i]);
        }
#pragma unroll 1
        for (unsigned g = 0; g < ngroups; ++g) {
            issue();
            cpa_wait<kAheadXY>();
            __syncwarp();  // the neighbour's copies have landed as well
            const uint32_t src = my0 + slot_r * kTileXY;
            uint32_t raw[16], nb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint4 v = lds128(src + r * kRowPitch);
                raw[4 * r] = v.x; raw[4 * r + 1] = v.y; raw[4 * r + 2] = v.z; raw[4 * r + 3] = v.w;
                nb[r] = lds32(src + r * kRowPitch + 16);
            }
            __syncwarp();  // everybody has read this slot before anybody's next copies overwrite it
            slot_r = (slot_r + 1 == kRingXY) ? 0 : slot_r + 1;
            if (MT != kMaskNone) dirty |= may_contain_masked<KIND, MT>(raw, mask_lo, mask_hi);
            // (two instantiations: in a full group every `present` test folds away)
            auto fold = [&](auto full, int nrows) {
                constexpr bool FULL = decltype(full)::value;
                uint32_t wb[16], wy[16], wx[16];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool present = FULL || r < nrows;
                    uint32_t c[5];
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[i] = present ? xf<KIND, XFORM>(raw[4 * r + i]) : 0u;
                    c[4] = xf<KIND, XFORM>(nb[r]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        wb[4 * r + i] = c[i];
                        wy[4 * r + i] = c[i] & prev[i];
                        wx[4 * r + i] = c[i] & c[i + 1];
                    }
                    if (present) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) prev[i] = c[i];
                    }
                }
                Bl.add16(wb, it);
                ABy.add16(wy, it);
                ABx.add16(wx, it);
                if (col_warp) {
                    uint32_t cw[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) cw[r] = role == 1u ? wb[4 * r] : role == 2u ? wb[4 * r + 3] : 0u;
                    Col.add_chunk(cw, cit);
                    ++cit;
                    if (cit == LC::kFlushChunks) flush_col();
                }
            };
            if (g < groups) fold(std::true_type{}, 4);
            else fold(std::false_type{}, tail_rows);
            ++it;
            if (it == L::kFlushChunks) flush(true);
        }
        slot_r = slot_w;  // the kAhead groups committed past the end of the item were empty
    }
    cpa_wait<0>();
    flush(true);
    flush_col();

    // one global atomic per non-zero counter and block
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned long long b = s_sum[0][lane], y = s_sum[1][lane], xx = s_sum[2][lane];
        // x axis: the first column is never a b, the last never an a
        if (s_sum[3][lane]) atomicAdd(&ws_x->minus.sums[S_B][lane], s_sum[3][lane]);
        if (s_sum[4][lane]) atomicAdd(&ws_x->minus.sums[S_A][lane], s_sum[4][lane]);
        // y axis: a-stream = b-stream shifted up one row (the launcher adds the first row, subtracts the last)
        if (b) { atomicAdd(&ws_y->plus.sums[S_A][lane], b); atomicAdd(&ws_y->plus.sums[S_B][lane], b); }
        if (y) atomicAdd(&ws_y->plus.sums[S_AB][lane], y);
        // x axis: every cur element is an a and a b but for the last / first column (subtracted by the launcher)
        if (b) { atomicAdd(&ws_x->plus.sums[S_A][lane], b); atomicAdd(&ws_x->plus.sums[S_B][lane], b); }
        if (xx) atomicAdd(&ws_x->plus.sums[S_AB][lane], xx);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&ws_y->plus.npairs, y_flat_pairs);
        atomicAdd(&ws_x->plus.npairs, x_pairs);
    }
    if (MT != kMaskNone) {
        if (__any_sync(0xFFFFFFFFu, dirty) && lane == 0) atomicOr(dirty_flag, 1u);
    }
}

template <int KIND, bool XFORM, int MT>
cudaError_t launch_xy(const CountJob& jy, Workspace* ws_y, Workspace* ws_x, unsigned int* dirty_flag, cudaStream_t stream) {
    auto kern = k_pairs_cols_xy<KIND, XFORM, MT>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static std::atomic<unsigned long long> attr_set{0ull};
    const unsigned long long dev_bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemXY);
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nx = jy.inner, rows = jy.outer * jy.n, pair_rows = rows - 1;
    const int64_t row_bytes = nx * 4, cpr = row_bytes / 16, spr = (cpr + 31) / 32;
    // segments: 4 to 8 rounds of items per warp, at least 64 rows each, least idle time in the last round
    const int64_t warps = (int64_t)sms * kWarpsXY;
    const int64_t max_seg = pair_rows / 64 > 0 ? pair_rows / 64 : 1;
    int64_t lo_seg = (4 * warps + spr - 1) / spr, hi_seg = (8 * warps) / spr;
    if (lo_seg < 1) lo_seg = 1;
    if (hi_seg < lo_seg) hi_seg = lo_seg;
    if (lo_seg > max_seg) lo_seg = max_seg;
    if (hi_seg > max_seg) hi_seg = max_seg;
    if (hi_seg - lo_seg > 4096) hi_seg = lo_seg + 4096;
    int64_t nseg = lo_seg;
    double best = -1.0;
    for (int64_t ns = lo_seg; ns <= hi_seg; ++ns) {
        const int64_t items = spr * ns, rounds = (items + warps - 1) / warps;
        const double eff = (double)items / (double)(rounds * warps);
        if (eff > best + 1e-9) { best = eff; nseg = ns; }
    }
    const int64_t seg_min = pair_rows / nseg, seg_extra = pair_rows - seg_min * nseg;
    int64_t blocks = (spr * nseg + kWarpsXY - 1) / kWarpsXY;
    if (blocks > sms) blocks = sms;
    kern<<<(unsigned)blocks, kThreadsXY, kSmemXY, stream>>>(
        static_cast<const char*>(jy.x), (long long)row_bytes, (long long)cpr, (long long)spr, (long long)nseg,
        (long long)seg_min, (long long)seg_extra, (uint32_t)jy.mask_bits, (uint32_t)(jy.mask_bits >> 32),
        (unsigned long long)(pair_rows * nx), (unsigned long long)(pair_rows * (nx - 1)), ws_y, ws_x, dirty_flag);
    note_fast_launch();
    return cudaGetLastError();
}

}  // namespace

bool pairs_xy_eligible(const CountJob& jy) {
    // jy: the job of the second-to-last axis (inner = length of the last axis)
    const int bytes = dtype_bytes(jy.dtype);
    if (bytes != 4) return false;
    if (jy.inner < 8 || (jy.inner * 4) % 16 != 0 || reinterpret_cast<uintptr_t>(jy.x) % 16 != 0) return false;
    if (jy.n < 2 || jy.outer * jy.n < 65) return false;  // too few rows to walk
    if (jy.mask_mode == kMaskNaN && !(jy.transform && jy.dtype == XBI_F32)) return false;
    if (jy.transform && jy.dtype != XBI_F32) return false;
    return true;
}

// Counts both axes into ws_y / ws_x (already cleared) up to the small follow-up passes, which it launches as well;
// what remains for the caller is one finish launch per axis.  *dirty_flag (cleared by the caller) is raised when a
// masked element was seen: both workspaces are then garbage and the axes must be recounted one by one.
cudaError_t launch_pairs_xy(const CountJob& jy, Workspace* ws_y, Workspace* ws_x, unsigned int* dirty_flag,
                            cudaStream_t stream) {
    cudaError_t e;
    const bool f32 = jy.dtype == XBI_F32 && jy.transform;
    if (f32) {
        if (jy.mask_mode == kMaskNone) e = launch_xy<kF32, true, kMaskNone>(jy, ws_y, ws_x, dirty_flag, stream);
        else if (jy.mask_mode == kMaskNaN) e = launch_xy<kF32, true, kMaskNaN>(jy, ws_y, ws_x, dirty_flag, stream);
        else e = launch_xy<kF32, true, kMaskValue>(jy, ws_y, ws_x, dirty_flag, stream);
    } else {
        if (jy.mask_mode == kMaskNone) e = launch_xy<kU32, false, kMaskNone>(jy, ws_y, ws_x, dirty_flag, stream);
        else e = launch_xy<kU32, false, kMaskValue>(jy, ws_y, ws_x, dirty_flag, stream);
    }
    if (e != cudaSuccess) return e;
    // y axis: first / last row planes and the block-boundary pairs, exactly as after the single-axis walkers
    // (unmasked formulation: the results only count when no masked element exists)
    CountJob jy_plain = jy;
    jy_plain.mask_mode = kMaskNone;
    e = launch_cols_followups(jy_plain, ws_y, stream);
    if (e != cudaSuccess) return e;
    // x axis: the job of the last axis over the same array
    const int64_t nx = jy.inner, rows = jy.outer * jy.n;
    CountJob jx = jy_plain;
    jx.outer = rows;
    jx.n = nx;
    jx.inner = 1;
    // the kernel's x pairs cover rows 1 .. rows-1; row 0: explicit three streams
    CountJob row0 = jx;
    row0.outer = 1;
    // (A_x = S - last column, B_x = S - first column over rows 1 .. rows-1: the kernel summed the two columns itself)
    return launch_pairs_generic_flat(row0, 0, nx - 1, &ws_x->plus, stream);
}

}  // namespace xbi
