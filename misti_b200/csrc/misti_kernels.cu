// misti_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of libmisti_b200.so.
//
// Two kernels make one batched evaluation:
//   misti_correct_kernel   one THREAD per item (four lanes per item for small batches): parameter mapping,
//                          negative-parameter test, the sequential coalescence-rate correction chain
//                          (CorrectLambdas / CorrectLambda / Smooth, incl. an iterate-faithful trust-region-reflective
//                          least-squares solver), the post-split closed-form coefficients, and the scalar pre-pass of
//                          the JSFS stage: the item's list of 128-byte segment records (one per interval with
//                          migration, one per RUN of intervals without).
//   misti_jsfs_kernel      one HALF WARP per item (3 chain states per lane, two items per warp, persistent grid):
//                          uniformised propagation + branch-length integrals on the 44-state chain for the segments
//                          with migration, the closed-form projector pass for the runs without, pulses,
//                          ancient-sample reset, collapse, closed-form one-population tail, the 7x44 / 7x8 JSFS
//                          contraction, normalisation, and -- fused -- the multinomial composite log-likelihood
//                          against every data row (bootstrap replicates).
//   misti_stiff_kernel     (rarely) the dense scaling-and-squaring step, FP64 MMA, for intervals too stiff to sweep,
//                          followed by the rest of that item's sweep and its results; returns at once when no item
//                          was parked.  Three launches per evaluation in all.
// There is no CPU path: every entry point below launches on the device or fails.
#include <cuda_runtime.h>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "misti_jsfs.cuh"
#include "misti_pair.cuh"
#include "misti_optim.cuh"

namespace {

using misti::ModelDesc;

constexpr int kCorrectThreads = 64;
constexpr int kCorrectMinBlocks = 8;
constexpr int kCoopMaxItems = 4096;  // up to here the four-lane variant of K1 still runs at about one warp per scheduler
constexpr int kJsfsWarps = 4;      // warps per block of the JSFS kernel (8 items per block)
constexpr int kJsfsMinBlocks = 3;  // occupancy target: caps the kernel at 168 registers per thread (12 warps per SM)
constexpr int kDeferPostMaxItems = 16384;  // inside the optimiser: up to here the post-split pass of cpfit mode runs in the JSFS kernel
constexpr int kLargeBatchItems = 6144;     // plain batches above this take the large-batch kernels (post-split pass with four lanes per
                                           // item, JSFS kernel with a pair of lanes per item); measured break-even: equal at 4 096 /
                                           // 6 144 items, 7 % faster at 8 192, 16 % at 16 384 (tools/midsize_probe.py)
constexpr int kPostSmemDoubles = 1024;  // shared-memory budget of the JSFS kernel for a model's post-split table (8 KB)
constexpr int kMaxChunk = 1 << 18;  // items per launch: the machine is full from 65 536 on; scratch is ~6 KB per item (records, rates)
constexpr int kPitch = 2;  // per interval and item the rate buffer holds la0, la1

static __device__ const double d_l8[8][8] = MISTI_L8_INIT;
static __device__ const unsigned char d_w8[7][8] = MISTI_W8_INIT;

// ------------------------------------------------------------------------------------------------
// K1: correction chain, one thread per item.  COOP = the variant for small batches (an optimiser step): FOUR lanes per
// item, all running the same chain on the same data, which share out the residual evaluations of a solver round
// (misti::eval_fj); results are bit-identical to the one-thread variant, the serial chain is shorter.
// ------------------------------------------------------------------------------------------------
// MODE 1 = the variant the on-device optimiser launches (item count and item list on the device, interruptible chains);
// MODE 2 = the diagnostic variant that records the per-interval solver trace (misti_eval_io.solve_trace); 0 = neither: the
// plain batched evaluation carries none of that code
template <int MINB, bool COOP, int MODE, int THREADS = kCorrectThreads>
__global__ void __launch_bounds__(THREADS, MINB)
misti_correct_kernel(int B, int P, const double* __restrict__ params, const int* __restrict__ model_ids, int model_default,
                     const ModelDesc* __restrict__ models, const double* __restrict__ times, const double* __restrict__ lh,
                     const double* __restrict__ gaux, const unsigned* __restrict__ cls_all, unsigned flags, double mixtureTH, const double* __restrict__ lc_inject, int numT_max, double* lc,
                     long stride, double* __restrict__ cpost, double* pr_out, int* __restrict__ status, int* __restrict__ nfev,
                     double* __restrict__ rec, int seg_cap, int* __restrict__ nseg, int* __restrict__ counters, int defer_post,
                     int n_models, int* __restrict__ solve_trace, const int* __restrict__ count_ptr, int regime,
                     const int* __restrict__ item_list, misti::ChainCkpt* __restrict__ ckpt, int* __restrict__ slice_ctl, int yield_below) {
    // count_ptr (nullable): the number of items lives on the device (the on-device optimiser packs the points of a round
    // behind a counter) and B is only the capacity the grid was sized for.  regime: 0 = run; 1 / 2 = this launch is one of
    // the pair (four lanes per item | one thread per item) of which only the variant that suits the round's size runs.
    constexpr bool FIT = MODE == 1;
    const bool defer_seg = (defer_post & 2) != 0;  // the segment pre-pass runs as a kernel of its own (misti_segments_kernel)
    defer_post &= 1;
    if (FIT && count_ptr) {
        const int n = *count_ptr;
        if ((regime == 1 && n > kCoopMaxItems) || (regime == 2 && n <= kCoopMaxItems)) return;
        B = n < B ? n : B;
    }
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid < 8) counters[gtid] = 0;  // work and park counters of the two kernels that follow in the stream
    // item_list (nullable): the launch covers the items item_list[0 .. B) -- the on-device optimiser keeps an item's scratch
    // (rates, records, checkpoint) at a fixed index while the list of items that run in a round is packed
    const int slot = COOP ? gtid >> 2 : gtid;
    const int b = (FIT && item_list && slot < B) ? item_list[slot] : slot;
    // one model for the whole batch (the usual case): its descriptor is staged in shared memory once per block
    __shared__ ModelDesc smd;
    if (!model_ids) {
        const int* src = reinterpret_cast<const int*>(models + model_default);
        int* dst = reinterpret_cast<int*>(&smd);
        for (int i = threadIdx.x; i < (int)(sizeof(ModelDesc) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    // One-block-per-SM launches (THREADS > kCorrectThreads, plain mode) with regime = 4: every thread of the block arrives
    // at barrier 1 exactly `align_total` times -- at the top of every interval of its chain (correct_lambdas_item), and here
    // for the intervals a chain never reached (items that fail, empty slots).  align_total = the block's longest chain.
    constexpr bool CAN_ALIGN = THREADS > kCorrectThreads && MODE == 0 && !COOP;
    int align_done = 0, align_total = 0;
    const bool aligning = CAN_ALIGN && (regime & 4) != 0;
    const int align_every = (regime >> 3) > 0 ? (regime >> 3) : 1;  // a barrier at every align_every-th interval
    if (aligning) {
        __shared__ int s_align, s_align_min;
        if (threadIdx.x == 0) { s_align = 0; s_align_min = INT_MAX; }
        __syncthreads();
        int mine = 0;
        if (slot < B) {
            const int mid = model_ids ? model_ids[b] : model_default;
            if ((unsigned)mid < (unsigned)n_models) mine = (models[mid].splitT + align_every - 1) / align_every;
        }
        atomicMax(&s_align, mine);
        if (mine > 0) atomicMin(&s_align_min, mine);
        __syncthreads();
        align_total = s_align;
        // chains of different length in one block: the short ones wait (at their make-up arrivals) for the long ones before
        // their post-split work.  In cpfit mode that work is deferred or cheap (a split-time grid: +3 %, run-away mixes: -20 %);
        // in the reference's default mode it is 87 trust-region solves per item and the two phases would add up (a split-time
        // grid without migration: 1.20 -> 1.63 ms), so there such a block runs without the barriers
        if (s_align_min != align_total && !(flags & MISTI_FLAG_CPFIT)) align_total = 0;
    }
    auto align_make_up = [&]() {
        if (aligning)
            for (; align_done < align_total; ++align_done) asm volatile("barrier.sync 1;" ::: "memory");
    };
    if (slot >= B) { align_make_up(); return; }
    if (model_ids && (unsigned)model_ids[b] >= (unsigned)n_models) {  // id -1: an empty slot of the on-device optimiser; any
        // other id outside the registered models (device-pointer calls are not validated on the host) is skipped as well
        nseg[b] = 0;
        status[b] = MISTI_SKIPPED;
        nfev[b] = 0;
        align_make_up();
        return;
    }
    const ModelDesc& md = model_ids ? models[model_ids[b]] : smd;
    const double* tt = times + md.grid_off;
    const double* ll = lh + 2 * md.grid_off;
    const double* ga = gaux + misti::kGridAux * md.grid_off;
    const unsigned* cls = cls_all + md.cls_off;
    const double* par = params + (long)b * P;
    double* lcb = lc + b;
    int st = MISTI_OK, nf = 0;
    double cp[3] = {0.0, 0.0, 0.0};
    bool cp_done = false;
    if (lc_inject) {
        for (int i = 0; i < md.n_params; ++i)
            if (par[i] < 0) st = MISTI_NEGATIVE_PARAM;
        const double* src = lc_inject + (long)b * 2 * numT_max;
        for (int t = 0; t < md.numT; ++t) {
            lcb[(kPitch * t) * stride] = src[2 * t];
            lcb[(kPitch * t + 1) * stride] = src[2 * t + 1];
        }
    } else {
        double* pr = pr_out ? pr_out + (long)b * (numT_max + 1) * 6 : nullptr;
        // defer_post (cpfit mode): the post-split pass is left to the lane groups of the JSFS kernel; they get
        // exp(nc1 - nc0) in the first coefficient slot
        double nc[2] = {0.0, 0.0};
        int* trace = nullptr;
        if (MODE == 2) {
            trace = solve_trace ? solve_trace + (long)b * 2 * numT_max : nullptr;
            if (trace && (!COOP || (gtid & 3) == 0))
                for (int t = 0; t < numT_max; ++t) { trace[2 * t] = 0; trace[2 * t + 1] = misti::kNoSolve; }
            if (COOP && (gtid & 3) != 0) trace = nullptr;  // the four lanes of an item hold the same values: one of them writes
        }
        if (FIT && ckpt) {
            // inside the on-device optimiser the chain may yield at an interval boundary when its time slice is used up
            // (ChainCkpt).  slice_ctl: [0] the time slice in microseconds (adapted by the optimiser between rounds),
            // [1] / [2] how many chains of this round ran to the end / were interrupted
            misti::ChainResume rs;
            rs.ck = ckpt + b;
            rs.budget_ns = b < yield_below ? 1000LL * slice_ctl[0] : 0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(rs.t_start));
            st = misti::correct_lambdas_item<COOP, true, false>(md, tt, ll, par, flags, mixtureTH, lcb, kPitch, stride, pr, &nf, ga, cp,
                                                                &cp_done, cls, defer_post ? nc : nullptr, nullptr, &rs);
            if (st == MISTI_PENDING) {  // interrupted: the checkpoint is written; the item runs on in the next round
                if (!COOP || (gtid & 3) == 0) { status[b] = MISTI_PENDING; atomicAdd(slice_ctl + 2, 1); }
                return;
            }
            if (!COOP || (gtid & 3) == 0) { ckpt[b].active = 0; atomicAdd(slice_ctl + 1, 1); }
        } else if (MODE == 2) {
            st = misti::correct_lambdas_item<COOP, false, true>(md, tt, ll, par, flags, mixtureTH, lcb, kPitch, stride, pr, &nf, ga, cp,
                                                                &cp_done, cls, defer_post ? nc : nullptr, trace, nullptr);
        } else {
            st = misti::correct_lambdas_item<COOP, false, false>(md, tt, ll, par, flags, mixtureTH, lcb, kPitch, stride, pr, &nf, ga, cp,
                                                                 &cp_done, cls, defer_post ? nc : nullptr, nullptr, nullptr,
                                                                 (aligning && align_total > 0) ? &align_done : nullptr, align_total, align_every);
        }
        if (defer_post && cp_done) cp[0] = exp(nc[1] - nc[0]);
    }
    int ns = 0;
    if (st == MISTI_OK) {
        if (!cp_done) misti::post_split_coeffs(md, tt, lcb, kPitch, stride, cp);
        // all per-interval scalar work of the JSFS stage: the item's segment records
        if (!defer_seg) st = misti::build_segments_item(md, tt, par, lcb, kPitch, stride, rec + (long)b * seg_cap * misti::kRecSlots, &ns, cls);
    }
    nseg[b] = ns;
    cpost[b] = cp[0];
    cpost[stride + b] = cp[1];
    cpost[2 * stride + b] = cp[2];
    status[b] = st;
    nfev[b] = nf;
    align_make_up();
}

// The segment pre-pass of a large batch as a kernel of its own: the same function on the same rates, one thread per item like
// the correction kernel -- but it is 14 % of that kernel's serial chain, needs a third of its registers and a few KB of code,
// so here the machine runs it with four times the warps per scheduler instead of at the end of a latency-bound chain.
__global__ void __launch_bounds__(128)
misti_segments_kernel(int B, int P, const double* __restrict__ params, const int* __restrict__ model_ids, int model_default,
                      const ModelDesc* __restrict__ models, const double* __restrict__ times, const unsigned* __restrict__ cls_all,
                      const double* __restrict__ lc, long stride, int* __restrict__ status, double* __restrict__ rec, int seg_cap,
                      int* __restrict__ nseg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || status[b] != MISTI_OK) return;  // (the correction kernel left nseg = 0 for the others)
    const ModelDesc& md = models[model_ids ? model_ids[b] : model_default];
    int ns = 0;
    const int st = misti::build_segments_item(md, times + md.grid_off, params + (long)b * P, lc + b, kPitch, stride,
                                              rec + (long)b * seg_cap * misti::kRecSlots, &ns, cls_all + md.cls_off);
    nseg[b] = ns;
    if (st != MISTI_OK) status[b] = st;
}

// Where the results of an item go (the optional pointers may be null)
struct ItemOut {
    const double* data;    // [R][8] data rows (7 counts, folded by the host if needed, + the likelihood constant)
    const double* data_t;  // the same transposed, [8][Rs]: lanes that stride the rows read consecutive words
    int R, Rs, unfolded;
    double *llh, *jafs, *jafs_raw;
    int *status, *terms;
    const int* row_ids;
    double* logs;          // [B][8] (nullable) the 7 logs of every item, for misti_score_rows_kernel (many data rows)
};

// Results of item b from its lane group: status, spectrum, and the fused composite likelihood over all data rows
// (bootstrap replicates; the lanes stride the rows).  Called after misti::jafs_finish, which left the logs in `ysm`.
__device__ __forceinline__ void emit_item(const ItemOut& o, const double* ysm, int lane, int b, int st, double raw_c, double jn_c,
                                          int nt, bool with_rows = true) {
    if (lane == 0) {
        o.status[b] = st;
        if (o.terms) o.terms[b] = nt;
    }
    if (lane < 7) {
        if (o.jafs) o.jafs[(long)b * 7 + lane] = st == MISTI_OK ? jn_c : nan("");
        if (o.jafs_raw) o.jafs_raw[(long)b * 7 + lane] = st == MISTI_OK ? raw_c : nan("");
    }
    if (!with_rows) {  // many data rows: misti_score_rows_kernel writes them from the item's logs
        if (lane < 8) o.logs[(long)b * 8 + lane] = lane < 7 ? ysm[misti::kTailLog + lane] : 0.0;
        return;
    }
    const double bad = (st == MISTI_NEGATIVE_PARAM || st == MISTI_CORRECTION_FAILED) ? -misti::kInf : nan("");
    const double* logj = ysm + misti::kTailLog;
    if (o.row_ids) {  // one data row per item
        const int r = o.row_ids[b];
        if (lane == 0) o.llh[b] = (st == MISTI_OK && r >= 0 && r < o.R) ? misti::score_row(o.data + 8 * (long)r, logj) : bad;
    } else if (st != MISTI_OK) {
        for (int r = lane; r < o.R; r += 16) o.llh[(long)b * o.R + r] = bad;
    } else {
        // every data row (bootstrap replicates): llh[r] = const_r + sum_i d_ri log p_i in the reference's order (score_row).
        // The stage is a stream of 8 bytes written per (item, row) pair; four rows per lane are in flight so that the
        // dependent multiply-adds of one row do not wait for the loads of the next.
        double lj[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) lj[c] = logj[c];
        const double* dt = o.data_t;
        const long Rs = o.Rs;
        double* dst = o.llh + (long)b * o.R;
        int r = lane;
        for (; r + 48 < o.R; r += 64) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = dt[7 * Rs + r + 16 * u];
#pragma unroll
            for (int c = 0; c < 7; ++c)
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] += dt[c * Rs + r + 16 * u] * lj[c];
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[r + 16 * u] = v[u];
        }
        for (; r < o.R; r += 16) {
            double v = dt[7 * Rs + r];
#pragma unroll
            for (int c = 0; c < 7; ++c) v += dt[c * Rs + r] * lj[c];
            dst[r] = v;
        }
    }
}

// MANY data rows (bootstrap replicates): the likelihoods llh[b][r] = const_r + sum_c d_rc log p_bc are a [B, 8] x [8, R]
// product whose cost is its OUTPUT (8 bytes per (item, row) pair), so the stage is a kernel of its own, shaped for the
// stores: the JSFS kernel leaves the 7 logs of an item in a 64-byte record, and here a warp keeps the data of 128 rows in
// registers (4 rows per lane, 8 doubles each), walks over its slice of the items -- 7 broadcast loads per item for 128
// pairs, against 8 loads per PAIR when every item streams the rows through L1 -- and writes 256 contiguous bytes per store
// instruction with a streaming hint.  Summation order = the reference's (score_row).
constexpr int kWarpRowsMin = 64;       // from this many data rows on
constexpr int kScoreItemsPerWarp = 64;
__global__ void __launch_bounds__(128)
misti_score_rows_kernel(int B, int R, const double* __restrict__ data_t, long Rs, const double* __restrict__ logs,
                        const int* __restrict__ status, double* __restrict__ llh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 128 + lane;
    const int i0 = (blockIdx.y * 4 + warp) * kScoreItemsPerWarp;
    if (i0 >= B) return;
    double d[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 8; ++c) d[u][c] = r0 + 32 * u < R ? data_t[c * Rs + r0 + 32 * u] : 0.0;
    const int i1 = i0 + kScoreItemsPerWarp < B ? i0 + kScoreItemsPerWarp : B;
    for (int b = i0; b < i1; ++b) {
        const int st = status[b];
        double* dst = llh + (long)b * R;
        if (st != MISTI_OK) {
            if (st == MISTI_SKIPPED) continue;
            const double bad = (st == MISTI_NEGATIVE_PARAM || st == MISTI_CORRECTION_FAILED) ? -misti::kInf : nan("");
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + 32 * u < R) __stcs(dst + r0 + 32 * u, bad);
            continue;
        }
        double lj[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) lj[c] = logs[(long)b * 8 + c];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double v = d[u][7];
#pragma unroll
            for (int c = 0; c < 7; ++c) v += d[u][c] * lj[c];
            if (r0 + 32 * u < R) __stcs(dst + r0 + 32 * u, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: expected JSFS + composite log-likelihood, one half warp per item
// ------------------------------------------------------------------------------------------------
template <int MINB, bool DEFER, bool FIT>
__global__ void __launch_bounds__(kJsfsWarps * 32, MINB)
misti_jsfs_kernel(int B, int P, const double* __restrict__ params, const int* __restrict__ model_ids, int model_default,
                  const ModelDesc* __restrict__ models, const double* __restrict__ rec, int seg_cap, const int* __restrict__ nseg,
                  long stride, const double* __restrict__ cpost, ItemOut out, misti::Cont* __restrict__ conts,
                  int* __restrict__ park_list, int* __restrict__ park_count, int* __restrict__ work_counter,
                  const double* __restrict__ post_tab, const double* __restrict__ lh, const int* __restrict__ count_ptr,
                  const int* __restrict__ item_list) {
    if (FIT && count_ptr) {  // the number of items lives on the device (see misti_correct_kernel)
        const int n = *count_ptr;
        B = n < B ? n : B;
        if (B <= 0) return;
    }
    // one item per 16-lane half warp (3 of the 44 chain states per lane), two items per warp
    __shared__ double ysm_all[kJsfsWarps * 2][misti::kGroupScratch];
    __shared__ misti::RunTable<misti::HalfWarpLanes> runtab;
    runtab.fill(threadIdx.x, blockDim.x);
    // one model for the whole batch (the usual case): its post-split table is staged in shared memory if it fits
    // (DEFER = the variant of the kernel that runs the post-split pass of cpfit mode, see eval_chunk)
    __shared__ double post_sm[DEFER ? kPostSmemDoubles : 1];
    const bool post_staged = DEFER && !model_ids &&
                             misti::kPostVals * models[model_default].post_per * misti::HalfWarpLanes::LANES <= kPostSmemDoubles;
    if (post_staged) {
        const int n = misti::kPostVals * models[model_default].post_per * misti::HalfWarpLanes::LANES;
        for (int i = threadIdx.x; i < n; i += blockDim.x) post_sm[i] = post_tab[models[model_default].post_off + i];
    }
    __syncthreads();
    const int half = threadIdx.x >> 4, lane = threadIdx.x & 15;
    double* ysm = ysm_all[half];
    const misti::HalfWarpLanes g;
    misti::LaneCtx<misti::HalfWarpLanes> L;
    L.init(g, ysm, &runtab);
    // The grid is persistent (as many blocks as fit on the device); every warp draws the next PAIR of items from a
    // counter, so the load balances itself although items differ in cost.  The two halves of a warp work on items
    // 2j and 2j + 1 in lock step; the odd one out re-reads the last item.
    while (true) {
        int i0 = 0;
        if ((threadIdx.x & 31) == 0) i0 = atomicAdd(work_counter, 2);
        i0 = __shfl_sync(0xffffffffu, i0, 0);
        if (i0 >= B) break;
        const bool has = i0 + (half & 1) < B;
        const int slot = has ? i0 + (half & 1) : B - 1;
        const int b = (FIT && item_list) ? item_list[slot] : slot;  // see misti_correct_kernel
        int st = out.status[b];
        // an item the correction kernel skipped (model id outside the registered models) runs along inactive
        const ModelDesc& md = models[st == MISTI_SKIPPED ? 0 : (model_ids ? model_ids[b] : model_default)];
        double raw_c, jn_c;
        int nt = 0;
        double cp[3] = {cpost[b], cpost[stride + b], cpost[2 * stride + b]};
        if (DEFER) {  // cpfit mode: cp[0] is exp(nc1 - nc0) of the correction chain, the coefficients are computed here
            const bool act = has && st == MISTI_OK && md.splitT < md.numT;
            const double* lh_last = lh + 2 * (long)(md.grid_off + md.numT - 1);
            if (post_staged) misti::post_split_cpfit_group(g, act, post_sm, md.post_per, lh_last, cp[0], cp);
            else misti::post_split_cpfit_group(g, act, post_tab + md.post_off, md.post_per, lh_last, cp[0], cp);
        }
        const int js = misti::jsfs_item<misti::HalfWarpLanes>(g, L, md, has && st == MISTI_OK, params + (long)b * P,
                                                              rec + (long)b * seg_cap * misti::kRecSlots, nseg[b], cp, &raw_c, &nt,
                                                              conts + b, false);
        const bool fin = misti::jafs_finish(g, ysm, &raw_c, out.unfolded != 0, &jn_c);  // all lanes, also of a group without an item
        const bool warp_rows = out.logs != nullptr;  // many data rows: scored by misti_score_rows_kernel from the item's logs
        if (has) {
            if (st == MISTI_OK) st = js;
            if (st == MISTI_STIFF) {  // parked at a stiff segment: misti_stiff_kernel takes the item from here and emits its results
                if (lane == 0) {
                    out.status[b] = MISTI_STIFF;
                    park_list[atomicAdd(park_count, 1)] = b;
                }
            } else {
                if (st == MISTI_OK && !fin) st = MISTI_NONFINITE;
                emit_item(out, ysm, lane, b, st, raw_c, jn_c, nt, !warp_rows);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 for large batches: expected JSFS + composite log-likelihood, one PAIR of lanes per item (16 items per warp), the
// chain's state, integrals and generator in registers (misti_pair.cuh).  Persistent grid; warps draw 16 items at a time.
// Items with a stiff segment or an infinite last interval are put on the redo list for the 16-lane kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kPairWarps = 4;
constexpr int kPairMinBlocks = 2;  // 255 registers per thread, 8 warps per SM
__global__ void __launch_bounds__(kPairWarps * 32, kPairMinBlocks)
misti_jsfs_pair_kernel(int B, int P, const double* __restrict__ params, const int* __restrict__ model_ids, int model_default,
                       const ModelDesc* __restrict__ models, const double* __restrict__ rec, int seg_cap, const int* __restrict__ nseg,
                       long stride, const double* __restrict__ cpost, ItemOut out, int* __restrict__ redo_list,
                       int* __restrict__ redo_count, int* __restrict__ work_counter) {
    __shared__ double ysm_all[kPairWarps * 16][48];
    const int lane = threadIdx.x & 31, role = lane & 1;
    double* ysm = ysm_all[threadIdx.x >> 1];
    while (true) {
        int i0 = 0;
        if (lane == 0) i0 = atomicAdd(work_counter, 16);
        i0 = __shfl_sync(0xffffffffu, i0, 0);
        if (i0 >= B) break;
        const int slot = i0 + (lane >> 1);
        const bool has = slot < B;
        const int b = has ? slot : B - 1;
        int st = out.status[b];
        // an item the correction kernel skipped (model id outside the registered models) runs along inactive
        const ModelDesc& md = models[st == MISTI_SKIPPED ? 0 : (model_ids ? model_ids[b] : model_default)];
        const double cp[3] = {cpost[b], cpost[stride + b], cpost[2 * stride + b]};
        // The 16 items of a warp run in lock step, segment by segment, which only pays when their segment lists agree in
        // length and in type (sweep / closed-form run) position by position -- the same model, or models that differ in split
        // time or rates only; and this kernel has no dense step for stiff segments or the infinite last interval.  So a warp
        // whose items disagree hands all of them, and any warp the items that hold such a segment, to the 16-lane kernel
        // (redo list) before any work is done on them.
        const double* recb = rec + (long)b * seg_cap * misti::kRecSlots;
        const bool act0 = has && st == MISTI_OK;
        const int ns = act0 ? nseg[b] : 0;
        bool redo = false;
        unsigned sig = 0;  // which segments are sweeps (positions folded modulo 32; the same model always agrees with itself)
#pragma unroll 4  // independent loads of lines the correction kernel wrote: in flight together
        for (int sg = role; sg < ns; sg += 2) {
            const int type = misti::seg_type(misti::seg_meta_bits(recb[(long)sg * misti::kRecSlots + 15]));
            redo |= type == misti::SEG_STIFF || type == misti::SEG_INF;
            if (type == misti::SEG_MIG) sig ^= 1u << (sg & 31);
        }
        redo = __shfl_xor_sync(0xffffffffu, (int)redo, 1) != 0 || redo;
        sig ^= __shfl_xor_sync(0xffffffffu, sig, 1);
        {
            const unsigned actm = __ballot_sync(0xffffffffu, act0 && !redo);
            const int ref = actm ? __ffs(actm) - 1 : 0;  // the first pair with work sets the pattern
            const unsigned sig0 = __shfl_sync(0xffffffffu, sig, ref);
            const int ns0 = __shfl_sync(0xffffffffu, ns, ref);
            if (__any_sync(0xffffffffu, act0 && !redo && (sig != sig0 || ns != ns0))) redo = true;
        }
        misti::PairResult res;
        misti::jsfs_pair_item(md, act0 && !redo, params + (long)b * P, recb, ns, cp, ysm, &res);
        res.redo = res.redo || (redo && act0);
        // normalise and take the logs (MigrationInference.py:583-613): the seven (unfolded) or four (folded) logs are split
        // between the two lanes and exchanged
        const bool unfolded = out.unfolded != 0;
        double tot = 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) tot += res.raw[c];
        double jn[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) jn[c] = res.raw[c] / tot;
        double a[4];
        if (unfolded) {
            a[0] = role ? jn[4] : jn[0]; a[1] = role ? jn[5] : jn[1]; a[2] = role ? jn[6] : jn[2]; a[3] = role ? 1.0 : jn[3];
        } else {
            a[0] = role ? jn[2] + jn[4] : jn[0] + jn[6]; a[1] = role ? jn[3] : jn[1] + jn[5]; a[2] = 1.0; a[3] = 1.0;
        }
        double lm[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) lm[u] = log(a[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) lo[u] = __shfl_xor_sync(0xffffffffu, lm[u], 1);
        double lj[7];
        if (unfolded) {
            lj[0] = role ? lo[0] : lm[0]; lj[1] = role ? lo[1] : lm[1]; lj[2] = role ? lo[2] : lm[2]; lj[3] = role ? lo[3] : lm[3];
            lj[4] = role ? lm[0] : lo[0]; lj[5] = role ? lm[1] : lo[1]; lj[6] = role ? lm[2] : lo[2];
        } else {
            lj[0] = role ? lo[0] : lm[0]; lj[1] = role ? lo[1] : lm[1]; lj[2] = role ? lm[0] : lo[0]; lj[3] = role ? lm[1] : lo[1];
            lj[4] = 0.0; lj[5] = 0.0; lj[6] = 0.0;
        }
        bool fin = true;
#pragma unroll
        for (int c = 0; c < 7; ++c) fin = fin && (fabs(lj[c]) <= DBL_MAX);
        // results (the counterpart of emit_item).  The 16 items of a warp are consecutive, so everything is staged through
        // the warp's shared memory and stored in contiguous runs (the outputs may be pinned host memory: 8-byte stores
        // scattered at a 56-byte stride cost a PCIe transaction each).  Items on the redo list are written too -- the
        // 16-lane kernel overwrites them later in the stream -- except their status, which it reads.
        const bool live = has && !res.redo;
        if (has && res.redo && role == 0) redo_list[atomicAdd(redo_count, 1)] = b;
        if (st == MISTI_OK && !fin) st = MISTI_NONFINITE;
        const int nv = B - i0 < 16 ? B - i0 : 16;
        {
            const int src = (lane & 15) * 2;
            const int st_i = __shfl_sync(0xffffffffu, st, src), nt_i = __shfl_sync(0xffffffffu, res.nterms, src);
            const bool live_i = __shfl_sync(0xffffffffu, (int)live, src) != 0;
            if (lane < 16 && live_i) {
                out.status[i0 + lane] = st_i;
                if (out.terms) out.terms[i0 + lane] = nt_i;
            }
        }
        double* wsm = ysm_all[(threadIdx.x >> 5) * 16];  // the warp's 768 doubles
        const int it16 = lane >> 1;
        auto flush = [&](double* dst, int per) {
            __syncwarp();
            for (int t = lane; t < nv * per; t += 32) dst[(long)i0 * per + t] = wsm[t];
            __syncwarp();
        };
        const bool ok = st == MISTI_OK;
        // lane 0 of a pair holds categories 0..3, lane 1 categories 4..6
        if (out.jafs) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * role + u < 7) wsm[it16 * 7 + 4 * role + u] = ok ? (role ? jn[(4 + u) % 7] : jn[u]) : nan("");
            flush(out.jafs, 7);
        }
        if (out.jafs_raw) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * role + u < 7) wsm[it16 * 7 + 4 * role + u] = ok ? (role ? res.raw[(4 + u) % 7] : res.raw[u]) : nan("");
            flush(out.jafs_raw, 7);
        }
        if (out.logs) {  // many data rows: misti_score_rows_kernel writes them from the item's logs
#pragma unroll
            for (int u = 0; u < 4; ++u) wsm[it16 * 8 + 4 * role + u] = role ? (u < 3 ? lj[(4 + u) % 7] : 0.0) : lj[u];
            flush(out.logs, 8);
            continue;
        }
        const double bad = (st == MISTI_NEGATIVE_PARAM || st == MISTI_CORRECTION_FAILED) ? -misti::kInf : nan("");
        if (out.row_ids) {  // one data row per item
            const int r = out.row_ids[b];
            if (role == 0) wsm[it16] = (ok && r >= 0 && r < out.R) ? misti::score_row(out.data + 8 * (long)r, lj) : bad;
            flush(out.llh, 1);
        } else if (out.R <= 32) {
            for (int r = role; r < out.R; r += 2) wsm[it16 * out.R + r] = ok ? misti::score_row(out.data + 8 * (long)r, lj) : bad;
            flush(out.llh, out.R);
        } else if (live) {
            for (int r = role; r < out.R; r += 2) out.llh[(long)b * out.R + r] = ok ? misti::score_row(out.data + 8 * (long)r, lj) : bad;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense step for stiff intervals: one BLOCK per parked item.  exp(M~ T) of the Van Loan augmented generator
// M~ = [[M, P0], [0, 0]] (45x45, padded to 48) by scaling and squaring of the UNIFORMISED matrix: for
// T' = T / 2^s with q T' <= 1/2 the series E = sum_k Pois(k; q T') A~^k (A~ = I + M~/q >= 0, sparse x dense
// products) is summed in shared memory, then squared s times with FP64 tensor-core MMAs (mma.sync m8n8k4.f64,
// "DMMA").  What is carried is F = E - I, never E: with rates of 1e7 next to rates of 1 the slow states have
// A~_cc = 1 - r/q with r/q ~ 1e-7, and storing "one minus tiny" costs the tiny part q/r ulps -- 5e-9 in the spectrum,
// measured against 50-digit values (tests/golden/stiff_exact.json), where the reference's own float64 result is good to
// 1e-15.  So the series runs on G = M~/q itself, D_k = A~^k - I = D_(k-1) + G D_(k-1) + G (D_0 = 0), F = sum_k Pois(k) D_k,
// and a squaring is F <- 2 F + F F.  Then P1 = P0 + F11 P0 and integralP = last column of F
// (MigrationInference.SolveDifEq, :530-540).
// After the stiff segment(s) the first warp of the block resumes the item's sweep (misti::jsfs_item from the continuation
// record) and emits its results; should the item meet another stiff segment, the block takes the dense step again.
// ------------------------------------------------------------------------------------------------
constexpr int kStiffThreads = 128;
constexpr int kStiffBlocksPerSm = 3;  // 59 KB of shared memory each
constexpr int kLd = 52;  // leading dimension of the 48x48 matrices in shared memory: with 52 = 4 mod 16 the 16 lanes of a
                         // half warp read 16 different bank pairs for both MMA operands (row gid, column tig and vice versa)

constexpr int kStiffSmem = (3 * 48 * kLd + 192 + 192 + 96) * (int)sizeof(double);

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C = 2 A + A * A for 48x48 (ld = kLd) matrices in shared memory: one squaring step of F = E - I, (I + F)^2 = I + (2 F + F F).
// Each of the 4 warps owns a 24x24 quadrant = 3 x 3 tiles of 8x8 and keeps its 9 accumulator tiles in registers, so a
// k-step loads 3 + 3 operand fragments for 9 MMAs.
__device__ void dense_square(const double* __restrict__ A, double* __restrict__ C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int r0 = 24 * (warp >> 1), c0 = 24 * (warp & 1);
    double acc[3][3][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            acc[i][j][0] = 2.0 * A[(r0 + 8 * i + gid) * kLd + c0 + 8 * j + 2 * tig];
            acc[i][j][1] = 2.0 * A[(r0 + 8 * i + gid) * kLd + c0 + 8 * j + 2 * tig + 1];
        }
#pragma unroll 4
    for (int kk = 0; kk < 12; ++kk) {
        double a[3], b[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) a[i] = A[(r0 + 8 * i + gid) * kLd + 4 * kk + tig];
#pragma unroll
        for (int j = 0; j < 3; ++j) b[j] = A[(4 * kk + tig) * kLd + c0 + 8 * j + gid];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            C[(r0 + 8 * i + gid) * kLd + c0 + 8 * j + 2 * tig] = acc[i][j][0];
            C[(r0 + 8 * i + gid) * kLd + c0 + 8 * j + 2 * tig + 1] = acc[i][j][1];
        }
}

__global__ void __launch_bounds__(kStiffThreads, kStiffBlocksPerSm)
misti_stiff_kernel(int P, const double* __restrict__ params, const int* __restrict__ model_ids, int model_default,
                   const ModelDesc* __restrict__ models, const double* __restrict__ times, const double* __restrict__ lc,
                   long stride, const double* __restrict__ rec, int seg_cap, const int* __restrict__ nseg,
                   const double* __restrict__ cpost, ItemOut out, misti::Cont* __restrict__ conts,
                   const int* __restrict__ item_list, const int* __restrict__ item_count, int defer_post,
                   const double* __restrict__ post_tab, const double* __restrict__ lh) {
    const int n_items = *item_count;
    if (n_items == 0) return;  // the usual case: nothing was parked
    extern __shared__ double sm[];
    double* E = sm;                      // [48][kLd]
    double* Y = sm + 48 * kLd;           // [48][kLd]
    double* Z = sm + 2 * 48 * kLd;       // [48][kLd]
    double* vec = sm + 3 * 48 * kLd;     // P[48], tmp[48], adiag[48], aug[48], coef[48][4]
    double* Pv = vec; double* tmp = vec + 48; double* adiag = vec + 96; double* aug = vec + 144; double* coef = vec + 192;
    int* colo = reinterpret_cast<int*>(vec + 384);  // [44][4] offsets of the generator's off-diagonal columns (col * kLd)
    for (int i = threadIdx.x; i < 44 * 4; i += kStiffThreads) colo[i] = misti::d_ell[i >> 2][i & 3].col * kLd;
    __syncthreads();
    __shared__ double s_scal[4];
    __shared__ int s_flag[3];
    // the resuming sweep runs in warp 0 (its first half warp owns the item, the second idles)
    __shared__ double s_ysm[2][misti::kGroupScratch];
    __shared__ misti::RunTable<misti::HalfWarpLanes> runtab;
    runtab.fill(threadIdx.x, blockDim.x);
    __syncthreads();
    const int tid = threadIdx.x;
    const misti::HalfWarpLanes g;
    misti::LaneCtx<misti::HalfWarpLanes> L;
    if (tid < 32) L.init(g, s_ysm[tid >> 4], &runtab);
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
        const int b = item_list[i];
        const ModelDesc& md = models[model_ids ? model_ids[b] : model_default];
        const double* par = params + (long)b * P;
        const double* tt = times + md.grid_off;
        misti::Cont* ct = conts + b;
        const double* rb = rec + (long)b * seg_cap * misti::kRecSlots;
        const int ns = nseg[b];
      for (int round = 0; round <= ns; ++round) {  // dense step(s), then resume; again if the sweep parks the item once more
        int seg = ct->seg;
        int nterms = ct->nterms;
        if (tid < 48) Pv[tid] = ct->P[tid];
        if (tid == 0) s_flag[2] = 0;
        __syncthreads();
        int st = MISTI_OK;
        while (seg < ns) {  // the stiff segment the item was parked at, and any that follow it directly
            const unsigned long long meta = misti::seg_meta_bits(rb[seg * misti::kRecSlots + 15]);
            if (misti::seg_type(meta) != misti::SEG_STIFF) break;  // an ordinary segment: back to the sweep kernel
            const int it = misti::seg_it(meta);
            const double la0 = lc[(kPitch * it) * stride + b], la1 = lc[(kPitch * it + 1) * stride + b];
            const double m0 = misti::band_rate(md, par, it, 0), m1 = misti::band_rate(md, par, it, 1);
            const double T = tt[it];
            const double rate[4] = {la0, la1, m0, m1};
            // q = max |M_cc|
            if (tid < 48) {
                double d = 0.0;
                if (tid < 44)
                    for (int k = 0; k < 4; ++k) d += (double)misti::d_diag[tid][k] * rate[k];
                tmp[tid] = d;
            }
            __syncthreads();
            if (tid == 0) {
                double q = 0.0;
                for (int r = 0; r < 44; ++r) q = tmp[r] > q ? tmp[r] : q;
                const bool bad = !(q > 0.0 && q <= DBL_MAX && T >= 0.0 && T <= DBL_MAX);
                s_scal[0] = bad ? 1.0 : q;
                s_flag[0] = bad ? 1 : 0;
            }
            __syncthreads();
            if (s_flag[0]) { st = MISTI_NONFINITE; break; }
            const double q = s_scal[0], qinv = 1.0 / q;
            // ancient-sample reset (TwoPopulations.py:246-262) and pulse (:361-377) of this interval
            if (it == md.sampleDate && it > 0) {
                if (tid == 0) {
                    double a2 = 0.0, a11 = 0.0;
                    for (int r = 0; r < 44; ++r) {
                        if (misti::d_anc2[r]) a2 += Pv[r];
                        if (misti::d_anc11[r]) a11 += Pv[r];
                    }
                    for (int r = 0; r < 44; ++r) Pv[r] = 0.0;
                    Pv[2] = a2; Pv[11] = a11;
                }
                __syncthreads();
            }
            if (md.n_pulses > 0) {
                const double pu0 = misti::pulse_rate(md, par, it, 0), pu1 = misti::pulse_rate(md, par, it, 1);
                if (pu0 + pu1 > 0) {
                    const double pr = pu0 + pu1;
                    const int src = pu0 > 0 ? 0 : 1;
                    const misti::PulseEntry* ent = src == 0 ? misti::d_pulse0 : misti::d_pulse1;
                    const unsigned char* rp = src == 0 ? misti::d_pulse0_rowptr : misti::d_pulse1_rowptr;
                    if (tid < 44) {
                        double acc = 0.0;
                        for (int e = rp[tid]; e < rp[tid + 1]; ++e) {
                            const misti::PulseEntry pe = ent[e];
                            acc += (double)pe.mult * pow(1.0 - pr, (double)pe.a) * pow(pr, (double)pe.b) * Pv[pe.col];
                        }
                        tmp[tid] = acc;
                    }
                    __syncthreads();
                    if (tid < 44) Pv[tid] = tmp[tid];
                    __syncthreads();
                }
            }
            // uniformised augmented generator
            if (tid < 48) {
                double d = 0.0;
                if (tid < 44)
                    for (int k = 0; k < 4; ++k) d += (double)misti::d_diag[tid][k] * rate[k];
                adiag[tid] = tid < 44 ? -d * qinv : 0.0;  // diagonal of G = M~ / q (NOT of A~ = I + G: see above)
                aug[tid] = tid < 44 ? Pv[tid] * qinv : 0.0;
                for (int e = 0; e < 4; ++e) {
                    double c = 0.0;
                    if (tid < 44) {
                        const misti::EllEntry en = misti::d_ell[tid][e];
                        c = (double)en.cnt * rate[en.kind] * qinv;
                    }
                    coef[tid * 4 + e] = c;
                }
            }
            // scaling: q T / 2^s <= 1/2
            const double qT = q * T;
            int sq = 0;
            double lam = qT;
            while (lam > 0.5) { lam *= 0.5; ++sq; }
            // F = sum_k Pois(k; lam) D_k,  D_k = A~^k - I = D_(k-1) + G D_(k-1) + G, D_0 = 0 (rows 44..47 of D stay zero)
            const double p0 = exp(-lam);
            for (int idx = tid; idx < 48 * kLd; idx += kStiffThreads) {
                Y[idx] = 0.0;
                Z[idx] = 0.0;
                E[idx] = 0.0;
            }
            __syncthreads();
            double p = p0, rr = lam;
            int k = 0;
            double* Ya = Y; double* Yb = Z;
            while (true) {
                ++k;
                p *= rr;
                rr = lam / (k + 1);
                // entry (r, c) = idx / 48, idx % 48 for idx = tid, tid + 128, ...: 128 = 2 * 48 + 32
                for (int r = tid / 48, c = tid % 48; r < 44; r += c + 32 >= 48 ? 3 : 2, c = c + 32 >= 48 ? c - 16 : c + 32) {
                    double acc = adiag[r] * Ya[r * kLd + c];
                    double gen = r == c ? adiag[r] : (c == 44 ? aug[r] : 0.0);  // G[r][c]
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc += coef[r * 4 + e] * Ya[colo[r * 4 + e] + c];
                        if (colo[r * 4 + e] == c * kLd) gen += coef[r * 4 + e];
                    }
                    const double d = Ya[r * kLd + c] + (acc + gen);
                    Yb[r * kLd + c] = d;
                    E[r * kLd + c] += p * d;
                }
                __syncthreads();
                double* t = Ya; Ya = Yb; Yb = t;
                if (rr < 1.0 && p < misti::kUnifTol * (1.0 - rr)) break;
                if (k > 200) { st = MISTI_NONFINITE; break; }
            }
            nterms += k;
            // squarings (DMMA): F <- 2 F + F F, sq times
            double* Ea = E; double* Eb = Ya;  // Ya is free now (Yb too)
            for (int j = 0; j < sq; ++j) {
                dense_square(Ea, Eb);
                __syncthreads();
                double* t = Ea; Ea = Eb; Eb = t;
            }
            // P1 = P0 + F11 P0, integral = F[:, 44]
            if (tid < 44) {
                double acc = 0.0;
                for (int c = 0; c < 44; ++c) acc += Ea[tid * kLd + c] * Pv[c];
                tmp[tid] = Pv[tid] + acc;
                const double I = Ea[tid * kLd + 44];
                if (it < md.sampleDate) ct->Ia[tid] += I;
                else ct->Ib[tid] += I;
            }
            __syncthreads();
            if (tid < 44) Pv[tid] = tmp[tid];
            __syncthreads();
            ++seg;
            if (st != MISTI_OK) break;
        }
        if (tid < 48) ct->P[tid] = tid < 44 ? Pv[tid] : 0.0;
        if (tid == 0) {
            ct->seg = seg;
            ct->nterms = nterms;
        }
        __syncthreads();  // the continuation record is complete (block-wide visibility of the global stores)
        if (tid < 32) {
            const int lane = tid & 15;
            const bool mine = tid < 16;
            double raw_c, jn_c;
            int nt = 0;
            double cp[3] = {cpost[b], cpost[stride + b], cpost[2 * stride + b]};
            if (defer_post)
                misti::post_split_cpfit_group(g, mine && st == MISTI_OK && md.splitT < md.numT, post_tab + md.post_off, md.post_per,
                                              lh + 2 * (long)(md.grid_off + md.numT - 1), cp[0], cp);
            int js = misti::jsfs_item<misti::HalfWarpLanes>(g, L, md, mine && st == MISTI_OK, par, rb, ns, cp, &raw_c, &nt, ct, true);
            const bool fin = misti::jafs_finish(g, s_ysm[tid >> 4], &raw_c, out.unfolded != 0, &jn_c);
            if (mine) {
                if (st != MISTI_OK) js = st;
                if (js == MISTI_STIFF && round < ns) {
                    if (lane == 0) s_flag[2] = 1;  // parked again at a later stiff segment
                } else {
                    if (js == MISTI_OK && !fin) js = MISTI_NONFINITE;
                    emit_item(out, s_ysm[0], lane, b, js, raw_c, jn_c, nt, out.logs == nullptr);
                }
            }
        }
        __syncthreads();
        if (!s_flag[2]) break;
        __syncthreads();
      }
    }
}

// likelihood tail alone: one warp per spectrum, lanes stride the data rows
__global__ void __launch_bounds__(128)
misti_score_kernel(int B, const double* __restrict__ spectra, const double* __restrict__ data, int R, int unfolded,
                   double* __restrict__ llh) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    double raw[7], jn[7], logj[7];
    for (int c = 0; c < 7; ++c) raw[c] = spectra[(long)b * 7 + c];
    const bool ok = misti::jafs_normalise_logs(raw, unfolded != 0, jn, logj);
    for (int r = lane; r < R; r += 32) llh[(long)b * R + r] = ok ? misti::score_row(data + 8 * (long)r, logj) : nan("");
}

// The post-split pass of cpfit mode as a kernel of its own (large batches): one half warp per item runs
// misti::post_split_cpfit_group -- 86 logs per item that would otherwise sit at the end of the correction kernel's serial
// chain (14 % of it), spread over all lanes of the machine.  In: cpost[b] = exp(nc1 - nc0) from the correction kernel; out:
// the three coefficients in cpost[0..2][b], and exp(nc1 - nc0) kept in cpost[3][b] for the rates-on-request path.  The same
// function, table and lane count as the variant inside the JSFS kernel (small batches): bit-identical coefficients.
__global__ void __launch_bounds__(128)
misti_post_split_kernel(int B, const int* __restrict__ model_ids, int model_default, const ModelDesc* __restrict__ models,
                        const int* __restrict__ status, long stride, double* __restrict__ cpost, const double* __restrict__ post_tab,
                        const double* __restrict__ lh, const int* __restrict__ count_ptr, const int* __restrict__ item_list) {
    if (count_ptr) {
        const int n = *count_ptr;
        B = n < B ? n : B;
    }
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool has = slot < B;
    const int b = has ? (item_list ? item_list[slot] : slot) : 0;
    const misti::HalfWarpLanes g;
    const int st = has ? status[b] : MISTI_SKIPPED;
    const ModelDesc& md = models[(st == MISTI_SKIPPED || !has) ? 0 : (model_ids ? model_ids[b] : model_default)];
    const bool act = has && st == MISTI_OK && md.splitT < md.numT;
    const double ed_in = has ? cpost[b] : 1.0;  // read by every lane before lane 0 overwrites the slot
    const double ed = act ? ed_in : 1.0;
    double cp[3];
    misti::post_split_cpfit_group(g, act, post_tab + md.post_off, md.post_per, lh + 2 * (long)(md.grid_off + md.numT - 1), ed, cp);
    g.sync();
    if (act && g.lane() < 3) cpost[g.lane() * stride + b] = cp[g.lane()];
    if (has && g.lane() == 3) cpost[3 * stride + b] = ed_in;
}

// The same pass for large PLAIN batches with FOUR lanes per item, each on a quarter of the intervals (consecutive slices,
// read straight from the grid's aux rows), joined by one scan of the survival factors and one sum over the four lanes.
// The 16-lane form spends two thirds of its 55.6 M warp instructions on slice bookkeeping, the scan and the reductions
// (5.4 intervals per lane against a fixed cost of ~600 instructions); one thread per item has none of that but is a serial
// chain of 86 logs at 3.5 warps per scheduler (0.055 ms).  Four lanes: 22 intervals per lane, 55 warps per SM.
__global__ void __launch_bounds__(128)
misti_post_split_quad_kernel(int B, const int* __restrict__ model_ids, int model_default, const ModelDesc* __restrict__ models,
                             const int* __restrict__ status, long stride, double* __restrict__ cpost, const double* __restrict__ times,
                             const double* __restrict__ gaux, const double* __restrict__ lh) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = gt >> 2, q = gt & 3;
    const bool has = slot < B;
    const int b = has ? slot : 0;
    const int st = has ? status[b] : MISTI_SKIPPED;
    const ModelDesc& md = models[(st == MISTI_SKIPPED || !has) ? 0 : (model_ids ? model_ids[b] : model_default)];
    const bool act = has && st == MISTI_OK && md.splitT < md.numT;
    const double ed_in = has ? cpost[b] : 1.0;  // read by every lane before lane 0 overwrites the slot
    const double ed = act ? ed_in : 1.0, wn = 1.0 / (1.0 + ed);
    const int n = act ? md.numT - 1 - md.splitT : 0, per = (n + 3) >> 2;
    const int t0 = md.splitT + q * per, t1 = t0 + per < md.splitT + n ? t0 + per : md.splitT + n;
    const double* tt = times + md.grid_off;
    const double* ga0 = gaux + misti::kGridAux * (long)md.grid_off;
    double c6 = 0, c3 = 0, c1 = 0, e1 = 1.0;  // relative to the start of this lane's slice
    for (int t = t0; t < t1; ++t) {
        const double T = tt[t];
        if (T == 0) continue;
        const double* ga = ga0 + misti::kGridAux * t;
        const double u = (ga[0] + ed * ga[1]) * wn;   // exp(-lam T), the fitted non-coalescence probability
        const double q1 = (ga[2] + ed * ga[3]) * wn;  // 1 - u, free of cancellation
        const double z = -log(u);
        const double il = z > 0 ? T / z : 0.0;  // 1 / lam
        const double e3 = e1 * e1 * e1;
        const double q3 = q1 * (1.0 + u + u * u), q6 = q3 * (1.0 + u * u * u);  // 1 - u^3, 1 - u^6
        c1 += z > 0 ? e1 * q1 * il : e1 * T;
        c3 += z > 0 ? e3 * q3 * (il * (1.0 / 3.0)) : e3 * T;
        c6 += z > 0 ? (e3 * e3) * q6 * (il * (1.0 / 6.0)) : (e3 * e3) * T;
        e1 *= u;
    }
    // survival factor at the start of the slice: product of e1 over the lower lanes of the quad
    double v = e1;
    {
        double t = __shfl_up_sync(0xffffffffu, v, 1, 4);
        if (q >= 1) v *= t;
        t = __shfl_up_sync(0xffffffffu, v, 2, 4);
        if (q >= 2) v *= t;
    }
    const double up = __shfl_up_sync(0xffffffffu, v, 1, 4);
    const double f1 = q == 0 ? 1.0 : up, f3 = f1 * f1 * f1;
    c1 *= f1; c3 *= f3; c6 *= f3 * f3;
    if (act && q == 3) {  // the infinite last interval
        const double* lh_last = lh + 2 * (long)(md.grid_off + md.numT - 1);
        const double lam = (1.0 + ed) / (1.0 / lh_last[0] + ed / lh_last[1]);
        const double il = 1.0 / lam, x1 = f1 * e1, x3 = x1 * x1 * x1;
        c1 += x1 * il; c3 += x3 * (il * (1.0 / 3.0)); c6 += (x3 * x3) * (il * (1.0 / 6.0));
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c3 += __shfl_xor_sync(0xffffffffu, c3, o);
        c6 += __shfl_xor_sync(0xffffffffu, c6, o);
    }
    if (act && q < 3) cpost[q * stride + b] = q == 0 ? c6 : (q == 1 ? c3 : c1);
    if (has && q == 3) cpost[3 * stride + b] = ed_in;  // kept for the rates-on-request path
}

// Reduction over the ITEMS of a batch, per data row: best[r] = max_b llh[b][r] and the item that attains it (the first one,
// as numpy.argmax) -- what the reference's bootstrap notebook computes from 9 009 result lines (test.bs/bs_conf_int.ipynb:
// per replicate the split time of the highest likelihood), done where the likelihoods are, so that a sweep returns 2 R
// numbers instead of B R.  Two steps: chunks of `per` items per thread (rows across the threads: coalesced), then the
// chunks in order, merged into the running best of the call (item_offset = first item of this launch in the call).
__global__ void __launch_bounds__(128)
misti_rowmax_partial_kernel(int B, int R, const double* __restrict__ llh, int per, double* __restrict__ pbest, int* __restrict__ pitem) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= R) return;
    const int b0 = c * per, b1 = b0 + per < B ? b0 + per : B;
    double best = llh[(long)b0 * R + r];
    int item = b0;
    for (int b = b0 + 1; b < b1; ++b) {
        const double v = llh[(long)b * R + r];
        if (v > best) { best = v; item = b; }
    }
    pbest[(long)c * R + r] = best;
    pitem[(long)c * R + r] = item;
}

__global__ void __launch_bounds__(128)
misti_rowmax_final_kernel(int nchunks, int R, const double* __restrict__ pbest, const int* __restrict__ pitem, int item_offset, int first,
                          double* __restrict__ best, int* __restrict__ item) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double bv = first ? pbest[r] : best[r];
    int bi = first ? pitem[r] + item_offset : item[r];
    for (int c = first ? 1 : 0; c < nchunks; ++c) {
        const double v = pbest[(long)c * R + r];
        if (v > bv) { bv = v; bi = pitem[(long)c * R + r] + item_offset; }
    }
    best[r] = bv;
    item[r] = bi;
}

// lc[(2t+g)*stride + b]  ->  out[b][numT_max][2].  With defer_post (cpfit mode) the correction kernel left the post-split
// rates out (nothing on the path needs them): they are computed here from exp(nc1 - nc0), one interval per thread.
__global__ void misti_gather_lc_kernel(int B, int numT_max, const int* __restrict__ model_ids, int model_default,
                                       const ModelDesc* __restrict__ models, const double* __restrict__ lc, long stride,
                                       double* __restrict__ out, int defer_post, const double* __restrict__ cpost,
                                       const double* __restrict__ times, const double* __restrict__ lh,
                                       const double* __restrict__ gaux, int n_models) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long n = (long)B * 2 * numT_max;
    if (i >= n) return;
    const int b = (int)(i / (2 * numT_max)), j = (int)(i % (2 * numT_max));
    const int mid = model_ids ? model_ids[b] : model_default;
    if ((unsigned)mid >= (unsigned)n_models) { out[i] = 0.0; return; }  // a skipped item
    const ModelDesc& md = models[mid];
    const int t = j >> 1;
    double v = 0.0;
    if (t < md.numT) {
        if (defer_post && t >= md.splitT)
            v = misti::post_split_cpfit_rate(md, t, times[md.grid_off + t], gaux + misti::kGridAux * (long)(md.grid_off + t),
                                             lh + 2 * (long)md.grid_off, cpost[b]);
        else
            v = lc[(kPitch * t + (j & 1)) * stride + b];
    }
    out[i] = v;
}

// ---- structure-table export kernels (TwoPopulations / OnePopulation mirror classes) ---------------
__global__ void misti_generator_kernel(int which, double l1, double l2, double m1, double m2, double* out) {
    const int r = threadIdx.x;
    if (which == 1) {
        if (r < 8)
            for (int c = 0; c < 8; ++c) out[r * 8 + c] = l1 * d_l8[r][c];
        return;
    }
    if (r >= 44) return;
    const double rate[4] = {l1, l2, m1, m2};
    for (int c = 0; c < 44; ++c) out[r * 44 + c] = 0.0;
    double d = 0.0;
    for (int k = 0; k < 4; ++k) d += (double)misti::d_diag[r][k] * rate[k];
    out[r * 44 + r] = -d;
    for (int e = 0; e < MISTI_ELL_WIDTH; ++e) {
        const misti::EllEntry en = misti::d_ell[r][e];
        if (en.cnt) out[r * 44 + en.col] += (double)en.cnt * rate[en.kind];
    }
}

__global__ void misti_pulse_kernel(const double* P0, double rate, int src, double* P1) {
    const int r = threadIdx.x;
    if (r >= 44) return;
    const misti::PulseEntry* ent = src == 0 ? misti::d_pulse0 : misti::d_pulse1;
    const unsigned char* rp = src == 0 ? misti::d_pulse0_rowptr : misti::d_pulse1_rowptr;
    double acc = 0.0;
    for (int e = rp[r]; e < rp[r + 1]; ++e) {
        const misti::PulseEntry pe = ent[e];
        acc += (double)pe.mult * pow(1.0 - rate, (double)pe.a) * pow(rate, (double)pe.b) * P0[pe.col];
    }
    P1[r] = acc;
}

__global__ void misti_ancient_kernel(const double* P0, double* P1) {
    const int r = threadIdx.x;
    if (r >= 44) return;
    double v = 0.0;
    if (r == 2) {
        for (int i = 0; i < 44; ++i)
            if (misti::d_anc2[i]) v += P0[i];
    } else if (r == 11) {
        for (int i = 0; i < 44; ++i)
            if (misti::d_anc11[i]) v += P0[i];
    }
    P1[r] = v;
}

__global__ void misti_state_to_jaf_kernel(int which, int* out) {
    const int r = threadIdx.x;
    if (which == 1) {
        if (r < 8)
            for (int c = 0; c < 7; ++c) out[r * 7 + c] = d_w8[c][r];
    } else if (r < 44) {
        for (int c = 0; c < 7; ++c) out[r * 7 + c] = misti::d_w44[c][r];
    }
}

// forward map true rates -> PSMC-apparent rates (CoalescentRates): one chain, one thread
__global__ void misti_coal_rates_kernel(const ModelDesc* __restrict__ models, int model, const double* __restrict__ times,
                                        const double* __restrict__ lh, const unsigned* __restrict__ cls_all,
                                        const double* __restrict__ params, double mu0, double mu1, double* __restrict__ lh_out,
                                        double* __restrict__ pr_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const ModelDesc& md = models[model];
    const double mu[2] = {mu0, mu1};
    misti::coalescent_rates_item(md, times + md.grid_off, lh + 2 * (long)md.grid_off, params, mu, cls_all + md.cls_off, lh_out, pr_out);
}

// ---- the optimisers on the device (misti_optim.cuh): one thread per simplex / walker -----------------
// A fit is a stream of ROUNDS, each the same launches with the same arguments (so a few rounds are captured once as a
// CUDA graph and replayed): propose -> correction kernel -> JSFS kernel -> stiff kernel -> apply.  Every simplex is a
// small state machine of its own; nothing waits for the slowest one:
//   * the points of a round are PACKED: a simplex that still runs reserves as many items as it has points behind a
//     device-side counter, so a round costs what its running simplices cost, and the evaluation kernels take the item
//     count from the device (no host round trip, no empty slots);
//   * look-ahead (two Nelder-Mead iterations per round, 4 (3N + 4) points per simplex) switches itself on when the
//     simplices still running are few enough for the round to stay in the flat part of the latency curve;
//   * basin-hopping walkers (BhConfig.niter >= 0) take their Metropolis decision and start their next local search in
//     the propose step of the round after their local search ended -- no barrier between the hops of different walkers.
struct FitState {
    double *sim, *fsim;          // [S][(N+1) N], [S][N+1]
    long long *iters, *fcalls;   // [S]
    int *status, *phase;         // [S]
    const int *model, *row;      // [S] the pair each simplex fits
    int *first, *cnt, *look;     // [S] the simplex's points of the current step: first item, how many; look-ahead used
    int *waiting;                // [S] some of them were interrupted in the correction kernel and run on (ChainCkpt)
    // walkers (null for plain fits)
    double *bh_x, *bh_best_x;    // [S][N]
    double *bh_energy, *bh_best_f, *bh_step;
    int *bh_ok, *bh_best_ok, *bh_done;
    long long *bh_nfev, *bh_fail, *bh_nstep, *bh_naccept, *bh_hop;
    misti::Pcg64* rng;
};

// device-side bookkeeping of a fit: one block of ints
enum { FC_ITEMS = 0,        // items packed in the current round (read by the evaluation kernels)
       FC_ROUND = 1,        // round counter
       FC_RUN0 = 2,         // simplices that submitted points, by round parity (this round's count decides the next one's look-ahead)
       FC_RUN1 = 3,
       FC_ROUNDS_USED = 4,  // rounds that evaluated something
       FC_LAST = 5,         // items of the last completed round (0 = every simplex has ended)
       FC_POINTS_LO = 6, FC_POINTS_HI = 7,  // all items evaluated so far (64 bits; an interrupted item counts once per slice)
       FC_LOOKBASE = 8,     // items handed out in the look-ahead region in the current round
       FC_SLICE_US = 9,     // time slice of the correction chains (ChainCkpt), adapted between rounds
       FC_SLICE_DONE = 10, FC_SLICE_PEND = 11,  // chains of the current round that ran to the end / were interrupted
       FC_SLICE_MIN = 12,
       FC_N = 16 };

constexpr int kFitMaxPts = 64 * 4 > 17 * 16 ? 64 * 4 : 17 * 16;  // doubles one simplex can submit per round

__global__ void misti_fit_propose_kernel(int S, misti::NmConfig cfg, misti::BhConfig bh, FitState st, double* __restrict__ params,
                                         int* __restrict__ item_model, int* __restrict__ item_row, int* __restrict__ item_list,
                                         const int* __restrict__ item_status, int* __restrict__ fc, int look_max_items, int cap_list) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int N = cfg.N;
    const int round = fc[FC_ROUND];
    const int slots_plain = misti::nm_slots(N, false);
    if (st.waiting[s]) {
        // some points of the simplex's current step were interrupted in the correction kernel (ChainCkpt): they run on, the
        // simplex waits -- only ITS step takes longer, the other simplices go on at their own pace
        int n = 0;
        for (int j = 0; j < st.cnt[s]; ++j) n += item_status[st.first[s] + j] == MISTI_PENDING;
        const int at = atomicAdd(fc + FC_ITEMS, n);
        if (at + n <= cap_list) {
            int k = 0;
            for (int j = 0; j < st.cnt[s]; ++j)
                if (item_status[st.first[s] + j] == MISTI_PENDING) item_list[at + k++] = st.first[s] + j;
        }
        atomicAdd(fc + FC_RUN0 + ((round + 1) & 1), 1);
        return;
    }
    double* sim = st.sim + (long)s * (N + 1) * N;
    double* fsim = st.fsim + (long)s * (N + 1);
    // look-ahead while the simplices that ran in the previous round are few (the count only ever falls)
    const int running_prev = round == 0 ? S : fc[FC_RUN0 + (round & 1)];
    misti::NmConfig c = cfg;
    c.lookahead = cfg.lookahead && N <= misti::kNmLookaheadMaxN && (long)running_prev * misti::nm_slots(N, true) <= look_max_items;
    c.slots = misti::nm_slots(N, c.lookahead != 0);
    double pts[kFitMaxPts];
    int n = 0;
    for (int pass = 0; pass < 2 && n == 0; ++pass) {
        // a walker whose local search has ended takes its Metropolis decision and starts the next one at once: a simplex
        // that is not finished submits points in EVERY round (the host ends the fit at the first round without points)
        if (bh.niter >= 0 && st.phase[s] == misti::NM_DONE && !st.bh_done[s]) {
            misti::BhWalker w;
            w.x = st.bh_x + (long)s * N; w.best_x = st.bh_best_x + (long)s * N;
            w.energy = st.bh_energy + s; w.best_f = st.bh_best_f + s; w.step = st.bh_step + s;
            w.ok = st.bh_ok + s; w.best_ok = st.bh_best_ok + s; w.done = st.bh_done + s;
            w.nfev = st.bh_nfev + s; w.failures = st.bh_fail + s; w.nstep = st.bh_nstep + s; w.naccept = st.bh_naccept + s;
            w.hop = st.bh_hop + s; w.rng = st.rng + s;
            misti::bh_advance(bh, N, w, sim, fsim, st.iters + s, st.fcalls + s, st.status + s, st.phase + s);
        }
        if (st.phase[s] == misti::NM_DONE) break;
        n = misti::nm_propose(c, sim, fsim, st.iters + s, st.fcalls + s, st.status + s, st.phase + s, pts);  // 0: the search has just ended
    }
    st.look[s] = c.lookahead;
    if (n <= 0) { st.first[s] = -1; st.cnt[s] = 0; return; }
    // where the points live: a simplex's own slots (fixed: an interrupted item keeps its scratch across rounds), or -- the
    // 4 (3N + 4) points of a look-ahead step -- a range of the shared look-ahead region behind them
    const int base = c.lookahead ? S * slots_plain + atomicAdd(fc + FC_LOOKBASE, n) : s * slots_plain;
    const int at = atomicAdd(fc + FC_ITEMS, n);
    if (at + n > cap_list || (c.lookahead && base + n > S * slots_plain + look_max_items)) { st.first[s] = -1; st.cnt[s] = 0; return; }  // cannot happen
    st.first[s] = base;
    st.cnt[s] = n;
    atomicAdd(fc + FC_RUN0 + ((round + 1) & 1), 1);
    const int model = st.model[s], row = st.row[s];
    for (int j = 0; j < n; ++j) {
        for (int k = 0; k < N; ++k) params[(long)(base + j) * N + k] = pts[j * N + k];
        item_model[base + j] = model;
        item_row[base + j] = row;
        item_list[at + j] = base + j;
    }
}

__global__ void misti_fit_apply_kernel(int S, misti::NmConfig cfg, FitState st, const double* __restrict__ params,
                                       const double* __restrict__ llh, const int* __restrict__ item_status, int* __restrict__ fc) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S && st.first[s] >= 0) {
        bool pending = false;
        for (int j = 0; j < st.cnt[s]; ++j) pending |= item_status[st.first[s] + j] == MISTI_PENDING;
        st.waiting[s] = pending ? 1 : 0;
        if (!pending) {
            const int N = cfg.N;
            misti::NmConfig c = cfg;
            c.lookahead = st.look[s];
            c.slots = misti::nm_slots(N, c.lookahead != 0);
            const long b0 = st.first[s];
            misti::nm_apply(c, st.sim + (long)s * (N + 1) * N, st.fsim + (long)s * (N + 1), st.iters + s, st.fcalls + s, st.status + s,
                            st.phase + s, params + b0 * N, llh + b0, true);  // the objective is -llh
        }
    }
    if (s == 0) {  // next round (the counters are not read by the other threads of this kernel)
        const int n = fc[FC_ITEMS], r = fc[FC_ROUND];
        fc[FC_LAST] = n;
        if (n > 0) {
            fc[FC_ROUNDS_USED] += 1;
            const unsigned lo = (unsigned)fc[FC_POINTS_LO], add = (unsigned)n;
            fc[FC_POINTS_LO] = (int)(lo + add);
            if (lo + add < lo) fc[FC_POINTS_HI] += 1;
        }
        // The time slice follows the work.  A chain that is cut advances by one slice per ROUND, and a round lasts at least as
        // long as its ordinary chains take: slicing trades the latency of the cut chains (and of their simplices) for that of
        // all the others.  With f = the share of this round's chains that were cut: most of them (f > 1/2) -- the typical chain
        // is longer than the slice, cutting only adds rounds: double it; otherwise minimum x (1 + 8 f): short while the
        // run-away chains are a few among thousands of ordinary ones, long when the fit is down to a few simplices and the
        // cut ones ARE the critical path.
        const int done = fc[FC_SLICE_DONE], pend = fc[FC_SLICE_PEND];
        if (done + pend > 0) {
            long long v;
            if (2 * pend > done + pend) v = 2LL * fc[FC_SLICE_US];
            else v = fc[FC_SLICE_MIN] + 8LL * fc[FC_SLICE_MIN] * pend / (done + pend);
            fc[FC_SLICE_US] = v < (1 << 24) ? (int)v : (1 << 24);
        }
        fc[FC_SLICE_DONE] = 0;
        fc[FC_SLICE_PEND] = 0;
        fc[FC_ITEMS] = 0;
        fc[FC_LOOKBASE] = 0;
        fc[FC_RUN0 + (r & 1)] = 0;  // read by this round's propose step; the next round counts into it
        fc[FC_ROUND] = r + 1;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
struct misti_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int sm_count = 148;
    // grids (pooled: grid g occupies intervals [grid_off[g], grid_off[g] + numT[g]); times padded with one 0)
    std::vector<int> grid_numT, grid_off;
    std::vector<double> h_times, h_lh, h_gaux;  // h_gaux: misti::grid_aux_row per interval
    int numT_max = 0;
    double *d_times = nullptr, *d_lh = nullptr, *d_gaux = nullptr;
    size_t d_grid_cap = 0;
    bool grids_dirty = false;
    // models
    std::vector<ModelDesc> h_models;
    std::vector<unsigned> h_cls;  // misti::interval_class of every interval of every model, pooled
    unsigned* d_cls = nullptr;
    size_t d_cls_cap = 0;
    std::vector<double> h_post;  // misti::post_split_table of every model, pooled
    double* d_post = nullptr;
    size_t d_post_cap = 0;
    ModelDesc* d_models = nullptr;
    size_t d_models_cap = 0;
    bool models_dirty = false;
    // data rows
    int R = 0, unfolded = 1;
    double* d_data = nullptr;
    size_t d_data_cap = 0;
    // batch buffers
    size_t cap = 0;
    int cap_numT = 0;
    int cap_seg = 0;
    double *d_lc = nullptr, *d_cpost = nullptr;
    double* d_rec = nullptr;  // segment records [cap][cap_seg][16]
    int* d_nseg = nullptr;
    int *d_status = nullptr, *d_nfev = nullptr;
    misti::Cont* d_conts = nullptr;       // continuation records of parked (stiff) items
    int *d_queue[2] = {nullptr, nullptr}; // item lists of the stiff rounds
    int* d_counts = nullptr;              // their lengths (one counter per round)
    // staging for host-pointer calls
    size_t st_cap = 0, st_capR = 0, st_capP = 0;
    int st_numT = 0;
    double *s_params = nullptr, *s_llh = nullptr, *s_jafs = nullptr, *s_jafs_raw = nullptr;
    int *s_model_ids = nullptr, *s_terms = nullptr, *s_row_ids = nullptr;
    double *s_lc_io = nullptr, *s_pr = nullptr;
    size_t s_lc_io_cap = 0, s_pr_cap = 0;
    int* s_trace = nullptr;
    size_t s_trace_cap = 0;
    double* d_logs = nullptr;    // [cap][8] per-item logs for misti_score_rows_kernel
    size_t d_logs_cap = 0;
    double* d_rowmax = nullptr;  // scratch of the per-row reduction over items: partial bests, then best[R]; items behind as ints
    size_t d_rowmax_cap = 0;
    unsigned char* d_nm = nullptr;  // state and batch buffers of misti_nelder_mead (one block, carved up per call)
    size_t d_nm_cap = 0;
    int* h_nm_counts = nullptr;     // pinned: copy of the fit's device-side counters (FC_*), refreshed after every batch of rounds
    cudaGraphExec_t nm_graph = nullptr;   // one round of misti_nelder_mead as a CUDA graph, kept while its arguments stay valid
    std::vector<unsigned long long> nm_graph_key;
    unsigned long long generation = 0;    // bumped whenever a device buffer moves or a launch argument of the kernels changes
    int nm_use_graph = 1;                 // tuning knob MISTI_NM_GRAPH
    int nm_rounds_per_graph = 4;          // rounds captured in one graph (tuning knob MISTI_NM_ROUNDS)
    int nm_look_max = kCoopMaxItems;      // look-ahead while a round of look-ahead steps stays below this many items (MISTI_NM_LOOK_MAX)
    int split_segments = -1;              // segment pre-pass as a kernel of its own (-1 = large plain batches in default mode; knob MISTI_SPLIT_SEGMENTS = 0 / 1)
    int post_quad = 1;                  // plain batches: the post-split kernel with four lanes per item (knob MISTI_POST_QUAD = 0: 16 lanes)
    int correct_big_blocks = 1;           // one-wave batches: one block per SM in the correction kernel (knob MISTI_CORRECT_BIG_BLOCKS = 0)
    int correct_align = 1;                // ... and a barrier at every interval of the chain (knob MISTI_CORRECT_ALIGN = 0 / k: none / at every k-th interval)
    int jsfs_pair = -1;                   // JSFS kernel with a pair of lanes per item (-1 = large batches; knob MISTI_JSFS_PAIR = 0 / 1)
    int score_kernel = 1;                 // many data rows: likelihood stage as a kernel of its own (knob MISTI_SCORE_KERNEL)
    int fit_slice_us = 200;               // time slice of a correction chain inside the on-device optimiser (MISTI_FIT_SLICE_US)
    bool fit_slice_forced = false;        // the knob was set: slices also in large sweeps
    int nm_graph_launches = 0;            // kernel launches per round of the kept graph
    cudaEvent_t nm_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double* d_score = nullptr;      // scratch of misti_score_spectra
    size_t d_score_cap = 0;
    double* d_small = nullptr;  // 44*44 + 2*44 doubles for the table export kernels
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;
    int64_t launches = 0;
    int jsfs_minb = kJsfsMinBlocks;
    int correct_minb = kCorrectMinBlocks;
    int correct_coop = -1;  // -1 = by batch size
    int max_chunk = kMaxChunk;  // items per launch (test knob MISTI_MAX_CHUNK: chunk boundaries with small batches)
    int nm_lookahead = -1;  // on-device Nelder-Mead: two iterations per round; -1 = by size (tuning knob MISTI_NM_LOOKAHEAD)
    int defer_post = -1;    // cpfit mode: post-split pass in the JSFS kernel; -1 = by batch size (tuning knob MISTI_DEFER_POST)
};

namespace {

int fail(misti_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, MISTI_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

template <class T>
int ensure(misti_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return 0;
    size_t ncap = *cap ? *cap : 1;
    while (ncap < need) ncap *= 2;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    CK(cudaMalloc((void**)p, ncap * sizeof(T)));
    *cap = ncap;
    ++ctx->generation;
    return 0;
}

template <class T>
int realloc_exact(misti_ctx* ctx, T** p, size_t n) {
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    if (n) CK(cudaMalloc((void**)p, n * sizeof(T)));
    ++ctx->generation;
    return 0;
}

int sync_tables(misti_ctx* ctx) {
    if (ctx->grids_dirty || ctx->models_dirty) ++ctx->generation;
    if (ctx->grids_dirty) {
        size_t need = ctx->h_times.size();
        if (need > ctx->d_grid_cap) {
            size_t ncap = ctx->d_grid_cap ? ctx->d_grid_cap : 256;
            while (ncap < need) ncap *= 2;
            if (ctx->d_times) CK(cudaFree(ctx->d_times));
            if (ctx->d_lh) CK(cudaFree(ctx->d_lh));
            if (ctx->d_gaux) CK(cudaFree(ctx->d_gaux));
            ctx->d_times = ctx->d_lh = ctx->d_gaux = nullptr;
            CK(cudaMalloc((void**)&ctx->d_times, ncap * sizeof(double)));
            CK(cudaMalloc((void**)&ctx->d_lh, 2 * ncap * sizeof(double)));
            CK(cudaMalloc((void**)&ctx->d_gaux, misti::kGridAux * ncap * sizeof(double)));
            ctx->d_grid_cap = ncap;
        }
        // the previous launches may still read the old tables: copies are stream-ordered behind them
        CK(cudaMemcpyAsync(ctx->d_times, ctx->h_times.data(), need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_lh, ctx->h_lh.data(), 2 * need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_gaux, ctx->h_gaux.data(), misti::kGridAux * need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->grids_dirty = false;
    }
    if (ctx->models_dirty) {
        size_t need = ctx->h_models.size();
        if (need > ctx->d_models_cap) {
            size_t ncap = ctx->d_models_cap ? ctx->d_models_cap : 16;
            while (ncap < need) ncap *= 2;
            if (ctx->d_models) CK(cudaFree(ctx->d_models));
            ctx->d_models = nullptr;
            CK(cudaMalloc((void**)&ctx->d_models, ncap * sizeof(ModelDesc)));
            ctx->d_models_cap = ncap;
        }
        CK(cudaMemcpyAsync(ctx->d_models, ctx->h_models.data(), need * sizeof(ModelDesc), cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->h_cls.size() > ctx->d_cls_cap) {
            size_t ncap = ctx->d_cls_cap ? ctx->d_cls_cap : 1024;
            while (ncap < ctx->h_cls.size()) ncap *= 2;
            if (ctx->d_cls) CK(cudaFree(ctx->d_cls));
            ctx->d_cls = nullptr;
            CK(cudaMalloc((void**)&ctx->d_cls, ncap * sizeof(unsigned)));
            ctx->d_cls_cap = ncap;
        }
        CK(cudaMemcpyAsync(ctx->d_cls, ctx->h_cls.data(), ctx->h_cls.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->h_post.size() > ctx->d_post_cap || !ctx->d_post) {
            size_t ncap = ctx->d_post_cap ? ctx->d_post_cap : 1024;
            while (ncap < ctx->h_post.size()) ncap *= 2;
            if (ctx->d_post) CK(cudaFree(ctx->d_post));
            ctx->d_post = nullptr;
            CK(cudaMalloc((void**)&ctx->d_post, ncap * sizeof(double)));
            ctx->d_post_cap = ncap;
        }
        if (!ctx->h_post.empty())
            CK(cudaMemcpyAsync(ctx->d_post, ctx->h_post.data(), ctx->h_post.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->models_dirty = false;
    }
    return 0;
}

// most two-population segments any registered model can have
int max_segments(const misti_ctx* ctx) {
    int m = 1;
    for (const ModelDesc& md : ctx->h_models) {
        const int n2 = md.splitT < md.numT ? md.splitT : md.numT;
        if (n2 > m) m = n2;
    }
    return m;
}

int ensure_batch(misti_ctx* ctx, size_t B) {
    const int seg_need = max_segments(ctx);
    if (B <= ctx->cap && ctx->cap_numT >= ctx->numT_max && ctx->cap_seg >= seg_need) return 0;
    size_t ncap = ctx->cap ? ctx->cap : 1024;
    while (ncap < B) ncap *= 2;
    int rc;
    CK(cudaStreamSynchronize(ctx->stream));  // the previous launches may still use the buffers
    if ((rc = realloc_exact(ctx, &ctx->d_lc, ncap * kPitch * (size_t)ctx->numT_max))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_rec, ncap * (size_t)seg_need * misti::kRecSlots))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_nseg, ncap))) return rc;
    ctx->cap_seg = seg_need;
    if ((rc = realloc_exact(ctx, &ctx->d_cpost, ncap * 4))) return rc;  // three coefficients + exp(nc1 - nc0) (misti_post_split_kernel)
    if ((rc = realloc_exact(ctx, &ctx->d_status, ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_nfev, ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_conts, ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_queue[0], ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->d_queue[1], ncap))) return rc;
    if (!ctx->d_counts && (rc = realloc_exact(ctx, &ctx->d_counts, (size_t)8))) return rc;
    ctx->cap = ncap;
    ctx->cap_numT = ctx->numT_max;
    return 0;
}

// device staging buffers of the host-pointer calls, for chunks of up to n items
int ensure_staging(misti_ctx* ctx, size_t n, size_t R, size_t Pe) {
    const int numT_max = ctx->numT_max;
    if (n <= ctx->st_cap && R <= ctx->st_capR && Pe <= ctx->st_capP && numT_max <= ctx->st_numT) return 0;
    int rc;
    size_t ncap = ctx->st_cap ? ctx->st_cap : 1024;
    while (ncap < n) ncap *= 2;
    const size_t nR = R > ctx->st_capR ? R : ctx->st_capR;
    const size_t nP = Pe > ctx->st_capP ? Pe : ctx->st_capP;
    if ((rc = realloc_exact(ctx, &ctx->s_params, ncap * nP))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_llh, ncap * nR))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_jafs, ncap * 7))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_jafs_raw, ncap * 7))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_model_ids, ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_terms, ncap))) return rc;
    if ((rc = realloc_exact(ctx, &ctx->s_row_ids, ncap))) return rc;
    ctx->st_cap = ncap; ctx->st_capR = nR; ctx->st_capP = nP; ctx->st_numT = numT_max;
    return 0;
}

}  // namespace

extern "C" {

int misti_abi_version(void) { return MISTI_ABI_VERSION; }

int misti_ctx_create(int device, void* stream, misti_ctx** out) {
    if (!out) return MISTI_E_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return MISTI_E_NODEV;
    misti_ctx* ctx = new (std::nothrow) misti_ctx();
    if (!ctx) return MISTI_E_ARG;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MISTI_E_NODEV; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MISTI_E_CUDA; }
        ctx->own_stream = true;
    }
    if (cudaFuncSetAttribute(misti_stiff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStiffSmem) != cudaSuccess) {
        delete ctx;
        return MISTI_E_CUDA;
    }
    if (const char* e = getenv("MISTI_JSFS_MINB")) ctx->jsfs_minb = atoi(e);
    if (const char* e = getenv("MISTI_CORRECT_MINB")) ctx->correct_minb = atoi(e);
    if (const char* e = getenv("MISTI_CORRECT_COOP")) ctx->correct_coop = atoi(e);
    if (const char* e = getenv("MISTI_DEFER_POST")) ctx->defer_post = atoi(e);
    if (const char* e = getenv("MISTI_NM_LOOKAHEAD")) ctx->nm_lookahead = atoi(e);
    if (const char* e = getenv("MISTI_NM_GRAPH")) ctx->nm_use_graph = atoi(e);
    if (const char* e = getenv("MISTI_SCORE_KERNEL")) ctx->score_kernel = atoi(e);
    if (const char* e = getenv("MISTI_JSFS_PAIR")) ctx->jsfs_pair = atoi(e);
    if (const char* e = getenv("MISTI_CORRECT_ALIGN")) { const int v = atoi(e); if (v >= 0 && v <= 64) ctx->correct_align = v; }
    if (const char* e = getenv("MISTI_CORRECT_BIG_BLOCKS")) ctx->correct_big_blocks = atoi(e);
    if (const char* e = getenv("MISTI_POST_QUAD")) ctx->post_quad = atoi(e);
    if (const char* e = getenv("MISTI_SPLIT_SEGMENTS")) ctx->split_segments = atoi(e);
    if (const char* e = getenv("MISTI_NM_LOOK_MAX")) { const int v = atoi(e); if (v >= 64 && v <= kMaxChunk / 2) ctx->nm_look_max = v; }
    if (const char* e = getenv("MISTI_FIT_SLICE_US")) { const int v = atoi(e); if (v >= 0) { ctx->fit_slice_us = v; ctx->fit_slice_forced = true; } }
    if (const char* e = getenv("MISTI_NM_ROUNDS")) { const int v = atoi(e); if (v >= 1 && v <= 64) ctx->nm_rounds_per_graph = v; }
    if (const char* e = getenv("MISTI_MAX_CHUNK")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kMaxChunk) ctx->max_chunk = v;
    }
    for (int i = 0; i < 3; ++i)
        if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { delete ctx; return MISTI_E_CUDA; }
    if (cudaMalloc((void**)&ctx->d_small, (44 * 44 + 2 * 44) * sizeof(double)) != cudaSuccess) { delete ctx; return MISTI_E_CUDA; }
    *out = ctx;
    return 0;
}

void misti_ctx_destroy(misti_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    void* ptrs[] = {ctx->d_times, ctx->d_lh, ctx->d_gaux, ctx->d_cls, ctx->d_post, ctx->d_models, ctx->d_data, ctx->d_lc, ctx->d_cpost, ctx->d_status, ctx->d_nfev, ctx->d_conts, ctx->d_queue[0], ctx->d_queue[1], ctx->d_counts,
                    ctx->d_rec, ctx->d_nseg, ctx->s_params, ctx->s_llh, ctx->s_jafs, ctx->s_jafs_raw, ctx->s_model_ids, ctx->s_terms, ctx->s_row_ids, ctx->s_lc_io,
                    ctx->s_pr, ctx->d_small, ctx->d_nm, ctx->d_score, ctx->s_trace, ctx->d_rowmax, ctx->d_logs};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 3; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 4; ++i)
        if (ctx->nm_ev[i]) cudaEventDestroy(ctx->nm_ev[i]);
    if (ctx->h_nm_counts) cudaFreeHost(ctx->h_nm_counts);
    if (ctx->nm_graph) cudaGraphExecDestroy(ctx->nm_graph);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* misti_last_error(const misti_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int misti_ctx_set_stream(misti_ctx* ctx, void* stream) {
    if (!ctx) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    ctx->ev_valid = false;
    ++ctx->generation;
    return 0;
}

int misti_ctx_synchronize(misti_ctx* ctx) {
    if (!ctx) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_ctx_reserve(misti_ctx* ctx, int32_t B, int32_t P, int32_t rows_per_item) {
    if (!ctx) return MISTI_E_ARG;
    if (B < 0 || P < 0 || P > MISTI_MAX_PARAMS || rows_per_item < 1)
        return fail(ctx, MISTI_E_ARG, "misti_ctx_reserve: bad arguments");
    if (B == 0 || ctx->h_models.empty()) return 0;
    CK(cudaSetDevice(ctx->device));
    const size_t n = B < ctx->max_chunk ? (size_t)B : (size_t)ctx->max_chunk;
    int rc;
    if ((rc = ensure_batch(ctx, n))) return rc;
    return ensure_staging(ctx, n, (size_t)rows_per_item, (size_t)(P > 0 ? P : 1));
}

int misti_add_grid(misti_ctx* ctx, int32_t numT, const double* times, const double* lh, int32_t* grid_id) {
    if (!ctx) return MISTI_E_ARG;
    if (numT < 1 || !lh || (numT > 1 && !times) || !grid_id) return fail(ctx, MISTI_E_ARG, "misti_add_grid: bad arguments");
    for (int i = 0; i < 2 * numT; ++i)
        if (!(lh[i] == lh[i])) return fail(ctx, MISTI_E_ARG, "misti_add_grid: NaN rate");
    const int off = (int)ctx->h_times.size();
    ctx->grid_numT.push_back(numT);
    ctx->grid_off.push_back(off);
    for (int i = 0; i < numT - 1; ++i) ctx->h_times.push_back(times[i]);
    ctx->h_times.push_back(0.0);  // padding: the last interval is infinite and has no length entry
    for (int i = 0; i < 2 * numT; ++i) ctx->h_lh.push_back(lh[i]);
    for (int i = 0; i < numT; ++i) {
        double row[misti::kGridAux];
        misti::grid_aux_row(lh + 2 * i, ctx->h_times[off + i], row);
        for (int j = 0; j < misti::kGridAux; ++j) ctx->h_gaux.push_back(row[j]);
    }
    if (numT > ctx->numT_max) ctx->numT_max = numT;
    ctx->grids_dirty = true;
    *grid_id = (int32_t)ctx->grid_numT.size() - 1;
    return 0;
}

int misti_add_model(misti_ctx* ctx, const misti_model_desc* d, int32_t* model_id) {
    if (!ctx) return MISTI_E_ARG;
    if (!d || !model_id) return fail(ctx, MISTI_E_ARG, "misti_add_model: null argument");
    if (d->grid_id < 0 || d->grid_id >= (int)ctx->grid_numT.size()) return fail(ctx, MISTI_E_ARG, "misti_add_model: unknown grid");
    const int numT = ctx->grid_numT[d->grid_id];
    if (d->split_t < 0 || d->split_t > numT) return fail(ctx, MISTI_E_ARG, "misti_add_model: split time outside the grid");
    if (d->sample_date < 0 || d->sample_date > d->split_t)
        return fail(ctx, MISTI_E_ARG, "misti_add_model: split time more recent than sample date");
    if (d->n_bands < 0 || d->n_bands > MISTI_MAX_BANDS || d->n_pulses < 0 || d->n_pulses > MISTI_MAX_PULSES ||
        d->n_params < 0 || d->n_params > MISTI_MAX_PARAMS)
        return fail(ctx, MISTI_E_ARG, "misti_add_model: too many bands / pulses / parameters");
    ModelDesc md;
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = d->split_t; md.sampleDate = d->sample_date;
    md.n_bands = d->n_bands; md.n_pulses = d->n_pulses; md.n_params = d->n_params;
    md.grid_off = ctx->grid_off[d->grid_id];
    for (int b = 0; b < d->n_bands; ++b) {
        if ((d->band_pop[b] != 0 && d->band_pop[b] != 1) || d->band_start[b] < 0 || d->band_end[b] <= d->band_start[b] ||
            d->band_end[b] > numT || d->band_opt[b] >= d->n_params || !(d->band_val[b] == d->band_val[b]))
            return fail(ctx, MISTI_E_ARG, "misti_add_model: invalid migration band");
        md.band_pop[b] = d->band_pop[b]; md.band_start[b] = d->band_start[b]; md.band_end[b] = d->band_end[b];
        md.band_opt[b] = d->band_opt[b] < 0 ? -1 : d->band_opt[b]; md.band_val[b] = d->band_val[b];
    }
    for (int b = 0; b < d->n_pulses; ++b) {
        if ((d->pulse_pop[b] != 0 && d->pulse_pop[b] != 1) || d->pulse_time[b] < 0 || d->pulse_time[b] >= numT ||
            d->pulse_opt[b] >= d->n_params || !(d->pulse_val[b] == d->pulse_val[b]))
            return fail(ctx, MISTI_E_ARG, "misti_add_model: invalid pulse");
        md.pulse_pop[b] = d->pulse_pop[b]; md.pulse_time[b] = d->pulse_time[b];
        md.pulse_opt[b] = d->pulse_opt[b] < 0 ? -1 : d->pulse_opt[b]; md.pulse_val[b] = d->pulse_val[b];
    }
    md.cls_off = (int)ctx->h_cls.size();
    for (int t = 0; t < numT; ++t) ctx->h_cls.push_back(misti::interval_class(md, t));
    md.post_off = (int)ctx->h_post.size();
    md.post_per = misti::post_split_per(numT, md.splitT, misti::HalfWarpLanes::LANES);
    ctx->h_post.resize(ctx->h_post.size() + (size_t)misti::kPostVals * md.post_per * misti::HalfWarpLanes::LANES);
    misti::post_split_table(numT, md.splitT, ctx->h_times.data() + md.grid_off, ctx->h_gaux.data() + (size_t)misti::kGridAux * md.grid_off,
                            misti::HalfWarpLanes::LANES, ctx->h_post.data() + md.post_off);
    ctx->h_models.push_back(md);
    ctx->models_dirty = true;
    *model_id = (int32_t)ctx->h_models.size() - 1;
    return 0;
}

int misti_clear_models(misti_ctx* ctx) {
    if (!ctx) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->grid_numT.clear(); ctx->grid_off.clear(); ctx->h_times.clear(); ctx->h_lh.clear(); ctx->h_gaux.clear(); ctx->h_models.clear(); ctx->h_cls.clear(); ctx->h_post.clear();
    ctx->numT_max = 0;
    ctx->grids_dirty = ctx->models_dirty = false;
    ++ctx->generation;
    return 0;
}

int misti_set_data(misti_ctx* ctx, int32_t R, const double* sfs, const double* llh_const, int32_t unfolded) {
    if (!ctx) return MISTI_E_ARG;
    if (R < 1 || !sfs) return fail(ctx, MISTI_E_ARG, "misti_set_data: need at least one data row");
    CK(cudaSetDevice(ctx->device));
    std::vector<double> rows((size_t)R * 16);  // [R][8] followed by the transposed copy [8][R]
    for (int r = 0; r < R; ++r) {
        const double* d = sfs + 8 * (size_t)r + 1;
        double* o = rows.data() + 8 * (size_t)r;
        double snps = 0.0;
        for (int i = 0; i < 7; ++i) snps += d[i];
        double c;
        if (unfolded) {
            for (int i = 0; i < 7; ++i) o[i] = d[i];
            c = lgamma(snps + 1.0);
            for (int i = 0; i < 7; ++i) c -= lgamma(d[i] + 1.0);
        } else {
            o[0] = d[0] + d[6]; o[1] = d[1] + d[5]; o[2] = d[2] + d[4]; o[3] = d[3];
            o[4] = o[5] = o[6] = 0.0;
            c = lgamma(snps + 1.0);
            c -= lgamma(o[0] + 1.0) + lgamma(o[1] + 1.0) + lgamma(o[2] + 1.0) + lgamma(o[3] + 1.0);
        }
        o[7] = llh_const ? llh_const[r] : c;
        for (int i = 0; i < 8; ++i) rows[8 * (size_t)R + (size_t)i * R + r] = o[i];
    }
    int rc;
    if ((rc = ensure(ctx, &ctx->d_data, &ctx->d_data_cap, rows.size()))) return rc;
    CK(cudaMemcpyAsync(ctx->d_data, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->R = R;
    ctx->unfolded = unfolded ? 1 : 0;
    ++ctx->generation;
    return 0;
}

static int eval_chunk(misti_ctx* ctx, int B, int P, const double* d_params, const int* d_model_ids, int model_default,
                      unsigned flags, double mixture_th, const double* d_lc_inject, double* d_llh, double* d_jafs,
                      double* d_jafs_raw, double* d_lc_out, double* d_pr, int* d_status_out, int* d_nfev_out, int* d_terms,
                      const int* d_row_ids, int* d_trace = nullptr, const int* d_count = nullptr, int defer_override = -1,
                      const int* d_item_list = nullptr, misti::ChainCkpt* d_ckpt = nullptr, int* d_slice_ctl = nullptr, int yield_below = 0) {
    int rc;
    if ((rc = ensure_batch(ctx, (size_t)B))) return rc;
    const long stride = (long)ctx->cap;
    const int numT_max = ctx->numT_max;
    // cpfit mode, small and medium batches: the post-split pass (one log per interval, independent intervals) runs in the
    // lanes of the JSFS kernel instead of at the end of the correction kernel's serial chain.  The work is the same either
    // way, so a full machine gains nothing (measured: 65 536 items 2 % slower), but an optimiser step does (1...1024 items
    // 0.33 -> 0.29 ms, 16 384 items 0.58 -> 0.54 ms).  The two variants differ in the order of summation (<= 1e-13).
    // Large batches: the pass is a kernel of its own between the two (misti_post_split_kernel; same numbers as the variant inside
    // the JSFS kernel, bit for bit).  defer_mode: 0 = in the correction chain (knob MISTI_DEFER_POST = 0), 1 = in the JSFS
    // kernel's lanes, 2 = kernel of its own.
    int defer_mode = 0;
    if ((flags & MISTI_FLAG_CPFIT) && !d_lc_inject) {
        const int knob = defer_override >= 0 ? defer_override : ctx->defer_post;
        defer_mode = knob < 0 ? (B <= (d_count ? kDeferPostMaxItems : kLargeBatchItems) ? 1 : 2) : knob;
        if (defer_mode < 0 || defer_mode > 2) defer_mode = 0;
    }
    const int defer_post = defer_mode != 0 ? 1 : 0;   // what the correction kernel and the rates-on-request path see
    // large plain batches in the reference's default mode: the segment pre-pass leaves the correction kernel's serial chain
    // too (misti_segments_kernel).  Measured at 65 536 items: default mode with migration 3.84 -> 3.53 ms; cpfit mode 0.604 ->
    // 0.624 ms (the chain is shorter by less than the kernel costs), so there the pre-pass stays where it is.
    const bool coop_k1 = ctx->correct_coop < 0 ? B <= kCoopMaxItems : ctx->correct_coop != 0;
    const bool split_seg = !d_count && !d_trace && !coop_k1 &&
                           (ctx->split_segments < 0 ? (B > kDeferPostMaxItems && !(flags & MISTI_FLAG_CPFIT)) : ctx->split_segments != 0);
    const int defer_k1 = defer_post | (split_seg ? 2 : 0);  // what the correction kernel is told
    const int defer_lanes = defer_mode == 1 ? 1 : 0;  // the JSFS / stiff kernels run the pass themselves
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    // small batches: four lanes per item (see misti_correct_kernel); the knob MISTI_CORRECT_COOP = 0 / 1 forces a variant.
    // With the item count on the device (d_count) both variants are launched and the one that suits the count runs.
    const bool coop = ctx->correct_coop < 0 ? B <= kCoopMaxItems : ctx->correct_coop != 0;
#define MISTI_LAUNCH_CORRECT2(MINB, COOP, NTHREADS, REGIME) MISTI_LAUNCH_CORRECT3(MINB, COOP, 0, NTHREADS, REGIME)
#define MISTI_LAUNCH_CORRECT3(MINB, COOP, FIT, NTHREADS, REGIME)                                                         \
    misti_correct_kernel<MINB, COOP, FIT><<<(unsigned)(((NTHREADS) + kCorrectThreads - 1) / kCorrectThreads), kCorrectThreads, 0, ctx->stream>>>( \
        B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_times, ctx->d_lh, ctx->d_gaux, ctx->d_cls, flags, mixture_th, \
        d_lc_inject, \
        numT_max, ctx->d_lc, stride, ctx->d_cpost, d_pr, ctx->d_status, ctx->d_nfev, ctx->d_rec, ctx->cap_seg, ctx->d_nseg, ctx->d_counts, \
        defer_k1, (int)ctx->h_models.size(), d_trace, d_count, REGIME, d_item_list, d_ckpt, d_slice_ctl, yield_below)
#define MISTI_LAUNCH_CORRECT(MINB)                                                                                       \
    if (coop) MISTI_LAUNCH_CORRECT2(MINB, true, 4L * B, 0); else MISTI_LAUNCH_CORRECT2(MINB, false, (long)B, 0)
    if (d_count) {  // the on-device optimiser: the pair of variants, of which the one that suits the round's item count runs
        if (ctx->correct_coop != 0)  // (the knob MISTI_CORRECT_COOP = 1 forces the four-lane variant for every round: whole capacity)
            MISTI_LAUNCH_CORRECT3(kCorrectMinBlocks, true, 1, 4L * ((ctx->correct_coop < 0 && B > kCoopMaxItems) ? kCoopMaxItems : B),
                                  ctx->correct_coop < 0 ? 1 : 0);
        if (ctx->correct_coop == 0 || (ctx->correct_coop < 0 && B > kCoopMaxItems))
            MISTI_LAUNCH_CORRECT3(kCorrectMinBlocks, false, 1, (long)B, ctx->correct_coop < 0 ? 2 : 0);
    } else if (d_trace) {  // diagnostics: the per-interval solver trace
        if (coop) MISTI_LAUNCH_CORRECT3(kCorrectMinBlocks, true, 2, 4L * B, 0); else MISTI_LAUNCH_CORRECT3(kCorrectMinBlocks, false, 2, (long)B, 0);
    } else
    if (!coop && ctx->correct_minb == kCorrectMinBlocks && ctx->correct_big_blocks &&
               (((long)B > (long)ctx->sm_count * 384 && (long)B <= (long)ctx->sm_count * 512) || (long)B >= (long)ctx->sm_count * 768)) {
        // A batch that fills the machine with one-thread-per-item warps (up to 16 per SM at 128 registers): one block per SM with
        // all of the SM's warps instead of seven or eight blocks of two (several waves of such blocks for larger batches).  The warps of an SM then start in the same cycle and stay close
        // to each other in the code, which is what the kernel is short of: it executes 78 KB of distinct code, the SM's
        // instruction cache hits 78 %, and the GPC-level instruction cache runs at 69 % of its peak request rate (ncu).  Measured at
        // 65 408 items: 0.579 -> 0.555 ms (a barrier at every interval on top: 0.543 ms, not taken: items that fail leave the chain
        // early).  Several waves gain more (131 072 items 1.139 -> 1.065 ms, 262 144 items 2.53 -> 2.08 ms), except between one and
        // one and a half waves, where the few blocks of the second wave each cost a whole chain (98 304 items 0.953 -> 1.005 ms:
        // small blocks there).  Same code per thread: results do not depend on the block size.
        // (Only block sizes whose register bound comes out at the 128 of the two-warp blocks: ptxas then generates the same
        // arithmetic and results do not depend on the size of the batch.  Blocks of 256 / 320 / 384 threads for smaller one-wave
        // batches were 8 % faster WITH the 255 / 204 / 170 registers their launch bounds allow (30 000 items 0.361 -> 0.333 ms) --
        // but that code rounds differently (solver evaluation counts of run-away items changed in a test) -- and 3 % slower
        // with 128 (eight warps per SM do not crowd the instruction cache).)
#define MISTI_LAUNCH_CORRECT_BIG(T)                                                                                       \
        misti_correct_kernel<1, false, 0, T><<<(unsigned)((B + (T) - 1) / (T)), T, 0, ctx->stream>>>(                      \
            B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_times, ctx->d_lh, ctx->d_gaux, ctx->d_cls, flags, mixture_th, \
            d_lc_inject, numT_max, ctx->d_lc, stride, ctx->d_cpost, d_pr, ctx->d_status, ctx->d_nfev, ctx->d_rec, ctx->cap_seg, ctx->d_nseg, \
            ctx->d_counts, defer_k1, (int)ctx->h_models.size(), d_trace, d_count, ctx->correct_align ? (4 | (ctx->correct_align << 3)) : 0, d_item_list, d_ckpt, d_slice_ctl, yield_below)
        if ((long)B <= (long)ctx->sm_count * 448) MISTI_LAUNCH_CORRECT_BIG(448);
        else MISTI_LAUNCH_CORRECT_BIG(512);
#undef MISTI_LAUNCH_CORRECT_BIG
    } else
    switch (ctx->correct_minb) {  // register budget per thread: 4 -> 255, 8 -> 128, 12 -> 80 (tuning knob MISTI_CORRECT_MINB)
        case 4: MISTI_LAUNCH_CORRECT(4); break;
        case 12: MISTI_LAUNCH_CORRECT(12); break;
        default: MISTI_LAUNCH_CORRECT(kCorrectMinBlocks); break;
    }
#undef MISTI_LAUNCH_CORRECT
#undef MISTI_LAUNCH_CORRECT2
#undef MISTI_LAUNCH_CORRECT3
    CK(cudaGetLastError());
    if (split_seg) {
        misti_segments_kernel<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(
            B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_times, ctx->d_cls, ctx->d_lc, stride, ctx->d_status, ctx->d_rec,
            ctx->cap_seg, ctx->d_nseg);
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    if (defer_mode == 2 && !d_count && ctx->post_quad) {
        misti_post_split_quad_kernel<<<(unsigned)((4L * B + 127) / 128), 128, 0, ctx->stream>>>(
            B, d_model_ids, model_default, ctx->d_models, ctx->d_status, stride, ctx->d_cpost, ctx->d_times, ctx->d_gaux, ctx->d_lh);
        CK(cudaGetLastError());
        ctx->launches += 1;
    } else if (defer_mode == 2) {
        misti_post_split_kernel<<<(unsigned)((16L * B + 127) / 128), 128, 0, ctx->stream>>>(
            B, d_model_ids, model_default, ctx->d_models, ctx->d_status, stride, ctx->d_cpost, ctx->d_post, ctx->d_lh, d_count, d_item_list);
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    int blocks = (B + 2 * kJsfsWarps - 1) / (2 * kJsfsWarps);
    const int per_sm = (ctx->jsfs_minb >= 2 && ctx->jsfs_minb <= 5) ? ctx->jsfs_minb : kJsfsMinBlocks;
    const int max_blocks = ctx->sm_count * per_sm;  // persistent grid: exactly the blocks that are resident together
    if (blocks > max_blocks) blocks = max_blocks;
    ItemOut out;
    out.data = ctx->d_data; out.data_t = ctx->d_data + 8 * (size_t)ctx->R; out.R = ctx->R; out.Rs = ctx->R; out.unfolded = ctx->unfolded;
    out.llh = d_llh; out.jafs = d_jafs; out.jafs_raw = d_jafs_raw; out.status = ctx->d_status; out.terms = d_terms;
    out.row_ids = d_row_ids;
    out.logs = nullptr;
    if (!d_row_ids && ctx->R >= kWarpRowsMin && ctx->score_kernel) {
        if ((rc = ensure(ctx, &ctx->d_logs, &ctx->d_logs_cap, (size_t)ctx->cap * 8))) return rc;
        out.logs = ctx->d_logs;
    }
    // Large plain batches: the pair-of-lanes kernel takes every item it can (all but those with a stiff segment or an
    // infinite last interval), the 16-lane kernel then runs over the redo list (usually empty: it returns at once).
    const bool pair_kernel = !d_count && !defer_lanes && (ctx->jsfs_pair < 0 ? B > kLargeBatchItems : ctx->jsfs_pair != 0);
    if (pair_kernel) {
        int pblocks = (B + 16 * kPairWarps - 1) / (16 * kPairWarps);
        if (pblocks > ctx->sm_count * kPairMinBlocks) pblocks = ctx->sm_count * kPairMinBlocks;
        misti_jsfs_pair_kernel<<<pblocks, kPairWarps * 32, 0, ctx->stream>>>(
            B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_rec, ctx->cap_seg, ctx->d_nseg, stride, ctx->d_cpost, out,
            ctx->d_queue[1], ctx->d_counts + 6, ctx->d_counts + 5);
        CK(cudaGetLastError());
        ctx->launches += 1;
        misti_jsfs_kernel<kJsfsMinBlocks, false, true><<<blocks, kJsfsWarps * 32, 0, ctx->stream>>>(
            B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_rec, ctx->cap_seg, ctx->d_nseg, stride, ctx->d_cpost,
            out, ctx->d_conts, ctx->d_queue[0], ctx->d_counts, ctx->d_counts + 4, ctx->d_post, ctx->d_lh, ctx->d_counts + 6, ctx->d_queue[1]);
    } else
#define MISTI_LAUNCH_JSFS(MINB)                                                                                            \
    if (defer_lanes) MISTI_LAUNCH_JSFS2(MINB, true); else MISTI_LAUNCH_JSFS2(MINB, false)
#define MISTI_LAUNCH_JSFS2(MINB, DEFER) if (d_count) MISTI_LAUNCH_JSFS3(kJsfsMinBlocks, DEFER, true); else MISTI_LAUNCH_JSFS3(MINB, DEFER, false)
#define MISTI_LAUNCH_JSFS3(MINB, DEFER, FIT)                                                                               \
    misti_jsfs_kernel<MINB, DEFER, FIT><<<blocks, kJsfsWarps * 32, 0, ctx->stream>>>(                                           \
        B, P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_rec, ctx->cap_seg, ctx->d_nseg, stride, ctx->d_cpost, \
        out, ctx->d_conts, ctx->d_queue[0], ctx->d_counts, ctx->d_counts + 4, ctx->d_post, ctx->d_lh, d_count, d_item_list)
    switch (ctx->jsfs_minb) {  // register budget: 2 -> 255, 3 -> 168, 4 -> 128, 5 -> 96 (tuning knob MISTI_JSFS_MINB)
        case 2: MISTI_LAUNCH_JSFS(2); break;
        case 4: MISTI_LAUNCH_JSFS(4); break;
        case 5: MISTI_LAUNCH_JSFS(5); break;
        default: MISTI_LAUNCH_JSFS(kJsfsMinBlocks); break;
    }
#undef MISTI_LAUNCH_JSFS
#undef MISTI_LAUNCH_JSFS2
#undef MISTI_LAUNCH_JSFS3
    CK(cudaGetLastError());
    // items parked at a stiff segment (usually none: the kernel then returns at once): dense scaling-and-squaring step,
    // rest of the sweep and results, one block per item
    misti_stiff_kernel<<<ctx->sm_count * kStiffBlocksPerSm, kStiffThreads, kStiffSmem, ctx->stream>>>(
        P, d_params, d_model_ids, model_default, ctx->d_models, ctx->d_times, ctx->d_lc, stride, ctx->d_rec, ctx->cap_seg,
        ctx->d_nseg, ctx->d_cpost, out, ctx->d_conts, ctx->d_queue[0], ctx->d_counts, defer_lanes, ctx->d_post, ctx->d_lh);
    CK(cudaGetLastError());
    ctx->launches += 2;
    if (out.logs) {  // many data rows: the likelihood stage as a kernel of its own, after every item has its logs
        const dim3 grid((unsigned)((ctx->R + 127) / 128), (unsigned)((B + 4 * kScoreItemsPerWarp - 1) / (4 * kScoreItemsPerWarp)));
        misti_score_rows_kernel<<<grid, 128, 0, ctx->stream>>>(B, ctx->R, out.data_t, (long)out.Rs, out.logs, ctx->d_status, d_llh);
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->ev_valid = true;
    ctx->launches += (d_count && ctx->correct_coop < 0 && B > kCoopMaxItems) ? 2 : 1;  // the correction kernel (or the pair of variants)
    if (d_lc_out) {
        const long n = (long)B * 2 * numT_max;
        misti_gather_lc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(B, numT_max, d_model_ids, model_default,
                                                                                     ctx->d_models, ctx->d_lc, stride, d_lc_out,
                                                                                     defer_post, ctx->d_cpost + (defer_mode == 2 ? 3 * stride : 0), ctx->d_times, ctx->d_lh,
                                                                                     ctx->d_gaux, (int)ctx->h_models.size());
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    if (d_status_out) CK(cudaMemcpyAsync(d_status_out, ctx->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_nfev_out) CK(cudaMemcpyAsync(d_nfev_out, ctx->d_nfev, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

// per-row reduction over the n items of one launch (likelihoods d_llh[n][R] on the device) into the running best of the call;
// layout of ctx->d_rowmax: partial bests [nchunks][R], best [R], then (as ints) partial items [nchunks][R], item [R]
constexpr int kRowmaxPer = 64;
static int rowmax_chunk(misti_ctx* ctx, int n, int R, const double* d_llh, long item_offset, bool first, double** d_best, int** d_item) {
    const int nchunks_max = (ctx->max_chunk + kRowmaxPer - 1) / kRowmaxPer;
    const size_t doubles = (size_t)(nchunks_max + 1) * R, ints = (size_t)(nchunks_max + 1) * R;
    int rc;
    if (first && (rc = ensure(ctx, &ctx->d_rowmax, &ctx->d_rowmax_cap, doubles + (ints + 1) / 2))) return rc;
    double* pbest = ctx->d_rowmax;
    double* best = pbest + (size_t)nchunks_max * R;
    int* pitem = reinterpret_cast<int*>(ctx->d_rowmax + doubles);
    int* item = pitem + (size_t)nchunks_max * R;
    const int nchunks = (n + kRowmaxPer - 1) / kRowmaxPer;
    const dim3 grid((unsigned)((R + 127) / 128), (unsigned)nchunks);
    misti_rowmax_partial_kernel<<<grid, 128, 0, ctx->stream>>>(n, R, d_llh, kRowmaxPer, pbest, pitem);
    CK(cudaGetLastError());
    misti_rowmax_final_kernel<<<(unsigned)((R + 127) / 128), 128, 0, ctx->stream>>>(nchunks, R, pbest, pitem, (int)item_offset, first ? 1 : 0, best, item);
    CK(cudaGetLastError());
    ctx->launches += 2;
    *d_best = best;
    *d_item = item;
    return 0;
}

int misti_eval_batch(misti_ctx* ctx, int32_t B, int32_t P, const double* params, const int32_t* model_ids, int32_t model_default,
                     uint32_t flags, double mixture_th, double* llh, const misti_eval_io* io) {
    if (!ctx) return MISTI_E_ARG;
    const bool want_rowmax = io && (io->row_best_llh || io->row_best_item);
    if (B < 0 || P < 0 || P > MISTI_MAX_PARAMS || (!llh && !want_rowmax)) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: bad arguments");
    if (want_rowmax && (io->row_ids || (flags & MISTI_FLAG_DEVICE_PTRS)))
        return fail(ctx, MISTI_E_ARG, "misti_eval_batch: row_best_* needs host pointers and no row_ids");
    if (B == 0) return 0;
    if (P > 0 && !params) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: params is null");
    if (ctx->h_models.empty()) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: no model registered");
    if (ctx->R < 1) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: no data rows (misti_set_data)");
    const int n_models = (int)ctx->h_models.size();
    if (!model_ids) {
        if (model_default < 0 || model_default >= n_models) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: unknown model");
        if (ctx->h_models[model_default].n_params > P) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: incorrect number of parameters");
    } else {
        for (const ModelDesc& md : ctx->h_models)
            if (md.n_params > P) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: incorrect number of parameters");
    }
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_tables(ctx))) return rc;
    misti_eval_io none;
    std::memset(&none, 0, sizeof(none));
    if (!io) io = &none;
    const int numT_max = ctx->numT_max, R = ctx->R;
    const int Rl = io->row_ids ? 1 : R;  // llh entries per item
    const int Pe = P > 0 ? P : 1;
    static const double dummy_param = 0.0;
    (void)dummy_param;

    if (flags & MISTI_FLAG_DEVICE_PTRS) {
        // asynchronous, everything already resident; chunks only bound the scratch size
        for (long off = 0; off < B; off += ctx->max_chunk) {
            const int n = (int)((B - off) < ctx->max_chunk ? (B - off) : ctx->max_chunk);
            rc = eval_chunk(ctx, n, P, params ? params + off * P : nullptr, model_ids ? model_ids + off : nullptr, model_default,
                            flags, mixture_th, io->lc_inject ? io->lc_inject + off * 2 * numT_max : nullptr, llh + off * Rl,
                            io->jafs ? io->jafs + off * 7 : nullptr, io->jafs_raw ? io->jafs_raw + off * 7 : nullptr,
                            io->lc_out ? io->lc_out + off * 2 * numT_max : nullptr,
                            io->pr_out ? io->pr_out + off * (numT_max + 1) * 6 : nullptr, io->status ? io->status + off : nullptr,
                            io->nfev ? io->nfev + off : nullptr, io->terms ? io->terms + off : nullptr,
                            io->row_ids ? io->row_ids + off : nullptr,
                            io->solve_trace ? io->solve_trace + off * 2 * numT_max : nullptr);
            if (rc) return rc;
        }
        return 0;
    }

    // host pointers: stage through context-owned device buffers, synchronous
    if (model_ids)
        for (int b = 0; b < B; ++b)
            if (model_ids[b] < 0 || model_ids[b] >= n_models) return fail(ctx, MISTI_E_ARG, "misti_eval_batch: unknown model id");
    for (long off = 0; off < B; off += ctx->max_chunk) {
        const int n = (int)((B - off) < ctx->max_chunk ? (B - off) : ctx->max_chunk);
        if ((rc = ensure_staging(ctx, (size_t)n, (size_t)Rl, (size_t)Pe))) return rc;  // Rl likelihoods per item
        const bool need_lc_io = io->lc_inject || io->lc_out;
        if (need_lc_io && (rc = ensure(ctx, &ctx->s_lc_io, &ctx->s_lc_io_cap, (size_t)n * 2 * numT_max))) return rc;
        if (io->pr_out && (rc = ensure(ctx, &ctx->s_pr, &ctx->s_pr_cap, (size_t)n * (numT_max + 1) * 6))) return rc;
        if (io->solve_trace && (rc = ensure(ctx, &ctx->s_trace, &ctx->s_trace_cap, (size_t)n * 2 * numT_max))) return rc;
        if (P > 0)
            CK(cudaMemcpyAsync(ctx->s_params, params + off * P, (size_t)n * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (model_ids)
            CK(cudaMemcpyAsync(ctx->s_model_ids, model_ids + off, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (io->row_ids)
            CK(cudaMemcpyAsync(ctx->s_row_ids, io->row_ids + off, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (io->lc_inject)
            CK(cudaMemcpyAsync(ctx->s_lc_io, io->lc_inject + off * 2 * numT_max, (size_t)n * 2 * numT_max * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
        if (io->pr_out) CK(cudaMemsetAsync(ctx->s_pr, 0, (size_t)n * (numT_max + 1) * 6 * sizeof(double), ctx->stream));
        // lc_inject and lc_out share one staging buffer: the inject copy is consumed by K1 before the gather writes
        rc = eval_chunk(ctx, n, P, ctx->s_params, model_ids ? ctx->s_model_ids : nullptr, model_default, flags, mixture_th,
                        io->lc_inject ? ctx->s_lc_io : nullptr, ctx->s_llh, ctx->s_jafs, ctx->s_jafs_raw,
                        io->lc_out ? ctx->s_lc_io : nullptr, io->pr_out ? ctx->s_pr : nullptr, nullptr, nullptr, ctx->s_terms,
                        io->row_ids ? ctx->s_row_ids : nullptr, io->solve_trace ? ctx->s_trace : nullptr);
        if (rc) return rc;
        if (llh) CK(cudaMemcpyAsync(llh + off * Rl, ctx->s_llh, (size_t)n * Rl * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (want_rowmax) {
            double* d_best = nullptr;
            int* d_item = nullptr;
            if ((rc = rowmax_chunk(ctx, n, R, ctx->s_llh, off, off == 0, &d_best, &d_item))) return rc;
            if (off + n >= B) {  // the last launch of the call: the running best is final
                if (io->row_best_llh) CK(cudaMemcpyAsync(io->row_best_llh, d_best, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                if (io->row_best_item) CK(cudaMemcpyAsync(io->row_best_item, d_item, (size_t)R * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            }
        }
        if (io->jafs) CK(cudaMemcpyAsync(io->jafs + off * 7, ctx->s_jafs, (size_t)n * 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (io->jafs_raw)
            CK(cudaMemcpyAsync(io->jafs_raw + off * 7, ctx->s_jafs_raw, (size_t)n * 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (io->lc_out)
            CK(cudaMemcpyAsync(io->lc_out + off * 2 * numT_max, ctx->s_lc_io, (size_t)n * 2 * numT_max * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
        if (io->pr_out)
            CK(cudaMemcpyAsync(io->pr_out + off * (numT_max + 1) * 6, ctx->s_pr, (size_t)n * (numT_max + 1) * 6 * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
        if (io->status) CK(cudaMemcpyAsync(io->status + off, ctx->d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (io->nfev) CK(cudaMemcpyAsync(io->nfev + off, ctx->d_nfev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (io->terms) CK(cudaMemcpyAsync(io->terms + off, ctx->s_terms, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (io->solve_trace)
            CK(cudaMemcpyAsync(io->solve_trace + off * 2 * numT_max, ctx->s_trace, (size_t)n * 2 * numT_max * sizeof(int),
                               cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int misti_fit(misti_ctx* ctx, int32_t S, int32_t N, const double* x0, const int32_t* model_ids, const int32_t* row_ids,
              uint32_t flags, double mixture_th, const misti_fit_opts* opts, misti_fit_result* res) {
    if (!ctx) return MISTI_E_ARG;
    if (S < 0 || N < 1 || N > MISTI_MAX_PARAMS || !x0 || !model_ids || !opts || !res || !res->x || !res->fun || !res->nit ||
        !res->nfev || !res->status)
        return fail(ctx, MISTI_E_ARG, "misti_fit: bad arguments");
    res->rounds = res->points = 0;
    res->graph = 0;
    if (S == 0) return 0;
    if (ctx->R < 1) return fail(ctx, MISTI_E_ARG, "misti_fit: no data rows (misti_set_data)");
    const bool walkers = opts->niter >= 0;
    if (walkers && (!opts->rng_state || opts->interval < 1 || !(opts->stepwise_factor > 0)))
        return fail(ctx, MISTI_E_ARG, "misti_fit: basin-hopping needs rng_state, interval >= 1 and a positive stepwise_factor");
    const int n_models = (int)ctx->h_models.size();
    for (int s = 0; s < S; ++s) {
        if (model_ids[s] < 0 || model_ids[s] >= n_models) return fail(ctx, MISTI_E_ARG, "misti_fit: unknown model id");
        if (ctx->h_models[model_ids[s]].n_params > N) return fail(ctx, MISTI_E_ARG, "misti_fit: incorrect number of parameters");
        if (row_ids && (row_ids[s] < 0 || row_ids[s] >= ctx->R)) return fail(ctx, MISTI_E_ARG, "misti_fit: unknown data row");
    }
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_tables(ctx))) return rc;
    misti::NmConfig cfg;
    cfg.N = N;
    // look-ahead is decided per round ON the device (few running simplices); the knob MISTI_NM_LOOKAHEAD = 0 forbids it
    cfg.lookahead = (ctx->nm_lookahead != 0 && N <= misti::kNmLookaheadMaxN) ? 1 : 0;
    cfg.slots = misti::nm_slots(N, false);
    cfg.xatol = opts->xatol; cfg.fatol = opts->fatol;
    cfg.maxiter = opts->maxiter < 0 ? LLONG_MAX : opts->maxiter;
    cfg.maxfev = opts->maxfev < 0 ? LLONG_MAX : opts->maxfev;
    misti::BhConfig bh;
    bh.niter = walkers ? opts->niter : -1;
    bh.interval = walkers ? opts->interval : 1;
    bh.beta = opts->T != 0 ? 1.0 / opts->T : misti::kInf;
    bh.target = opts->target_accept_rate; bh.factor = opts->stepwise_factor; bh.stepsize0 = opts->stepsize;
    // items: every simplex has its own nm_slots(N) slots (fixed, so that an interrupted item keeps its scratch across
    // rounds); behind them the region the look-ahead steps of the few simplices of a late round share
    int look_max = ctx->nm_look_max;
    const int stable_items = S * misti::nm_slots(N, false);
    if (stable_items > ctx->max_chunk) return fail(ctx, MISTI_E_ARG, "misti_fit: too many simplices for one call");
    if (cfg.lookahead && (long)stable_items + look_max > ctx->max_chunk) {  // a launch holds at most max_chunk items: the shared
        look_max = ctx->max_chunk - stable_items;                           // look-ahead region takes what the simplices' own slots leave
        if (look_max < misti::nm_slots(N, true)) { look_max = 0; cfg.lookahead = 0; }
    }
    const long cap = (long)stable_items + (cfg.lookahead ? look_max : 0);
    const int B = (int)cap;
    // Time slice of the correction chains inside a round (tuning knob MISTI_FIT_SLICE_US; 0 = chains are never interrupted).
    // Walkers and small sweeps run in the latency regime, where one run-away chain (ten times the ordinary one) would hold up
    // every simplex of its round: measured on 1 024 walkers x 20 hops, 9.6 s without, 5.6 s with a slice of 150 ... 350 us.
    // A large sweep of like fits (config 5b: 36 036 chains per round, all equally long) gains nothing and pays for the extra
    // rounds and the checkpoint code (0.146 -> 0.175 s): there the chains run through.
    const bool slicing = ctx->fit_slice_us > 0 && (walkers || stable_items <= kCoopMaxItems || ctx->fit_slice_forced);
    const long long budget_ns = slicing ? (long long)ctx->fit_slice_us * 1000 : 0;
    // one block of device memory, carved up (8-byte items first)
    const size_t n_sim = (size_t)S * (N + 1) * N, n_fsim = (size_t)S * (N + 1);
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_sim = carve(n_sim * 8), o_fsim = carve(n_fsim * 8), o_it = carve((size_t)S * 8), o_fc = carve((size_t)S * 8),
                 o_par = carve((size_t)B * N * 8), o_llh = carve((size_t)B * 8),
                 o_bx = carve((size_t)S * N * 8), o_bbx = carve((size_t)S * N * 8), o_be = carve((size_t)S * 8),
                 o_bbf = carve((size_t)S * 8), o_bst = carve((size_t)S * 8), o_bnf = carve((size_t)S * 8), o_bfl = carve((size_t)S * 8),
                 o_bns = carve((size_t)S * 8), o_bna = carve((size_t)S * 8), o_bhp = carve((size_t)S * 8), o_rng = carve((size_t)S * 32),
                 o_st = carve((size_t)S * 4), o_ph = carve((size_t)S * 4), o_mod = carve((size_t)S * 4), o_row = carve((size_t)S * 4),
                 o_first = carve((size_t)S * 4), o_look = carve((size_t)S * 4), o_bok = carve((size_t)S * 4), o_bbok = carve((size_t)S * 4),
                 o_bdn = carve((size_t)S * 4), o_bm = carve((size_t)B * 4), o_br = carve((size_t)B * 4), o_cnt = carve(FC_N * 4),
                 o_ncnt = carve((size_t)S * 4), o_wait = carve((size_t)S * 4), o_list = carve((size_t)B * 4),
                 o_ck = carve((size_t)B * sizeof(misti::ChainCkpt));
    if ((rc = ensure(ctx, &ctx->d_nm, &ctx->d_nm_cap, off))) return rc;
    if (!ctx->h_nm_counts) CK(cudaMallocHost((void**)&ctx->h_nm_counts, 4 * FC_N * sizeof(int)));
    for (int i = 0; i < 4; ++i)
        if (!ctx->nm_ev[i]) CK(cudaEventCreateWithFlags(&ctx->nm_ev[i], cudaEventDisableTiming));
    unsigned char* base = ctx->d_nm;
    FitState st;
    st.sim = (double*)(base + o_sim); st.fsim = (double*)(base + o_fsim);
    st.iters = (long long*)(base + o_it); st.fcalls = (long long*)(base + o_fc);
    st.status = (int*)(base + o_st); st.phase = (int*)(base + o_ph);
    int* d_mod = (int*)(base + o_mod); int* d_row = (int*)(base + o_row);
    st.model = d_mod; st.row = d_row;
    st.first = (int*)(base + o_first); st.look = (int*)(base + o_look);
    st.cnt = (int*)(base + o_ncnt); st.waiting = (int*)(base + o_wait);
    int* d_list = (int*)(base + o_list);
    misti::ChainCkpt* d_ck = budget_ns > 0 ? (misti::ChainCkpt*)(base + o_ck) : nullptr;
    st.bh_x = (double*)(base + o_bx); st.bh_best_x = (double*)(base + o_bbx); st.bh_energy = (double*)(base + o_be);
    st.bh_best_f = (double*)(base + o_bbf); st.bh_step = (double*)(base + o_bst);
    st.bh_ok = (int*)(base + o_bok); st.bh_best_ok = (int*)(base + o_bbok); st.bh_done = (int*)(base + o_bdn);
    st.bh_nfev = (long long*)(base + o_bnf); st.bh_fail = (long long*)(base + o_bfl); st.bh_nstep = (long long*)(base + o_bns);
    st.bh_naccept = (long long*)(base + o_bna); st.bh_hop = (long long*)(base + o_bhp);
    st.rng = (misti::Pcg64*)(base + o_rng);
    double* d_par = (double*)(base + o_par); double* d_llh = (double*)(base + o_llh);
    int* d_bm = (int*)(base + o_bm); int* d_br = (int*)(base + o_br); int* d_fc = (int*)(base + o_cnt);
    cudaStream_t sm = ctx->stream;
    // initial state: vertex 0 = x0, everything else zero (phase NM_INIT = 0, status -1 set below)
    CK(cudaMemsetAsync(base, 0, off, sm));
    CK(cudaMemcpy2DAsync(st.sim, (size_t)(N + 1) * N * 8, x0, (size_t)N * 8, (size_t)N * 8, S, cudaMemcpyHostToDevice, sm));
    CK(cudaMemsetAsync(st.status, 0xff, (size_t)S * 4, sm));
    CK(cudaMemcpyAsync(d_mod, model_ids, (size_t)S * 4, cudaMemcpyHostToDevice, sm));
    if (row_ids) CK(cudaMemcpyAsync(d_row, row_ids, (size_t)S * 4, cudaMemcpyHostToDevice, sm));
    if (walkers) {
        static_assert(sizeof(misti::Pcg64) == 32, "four 64-bit words per generator");
        CK(cudaMemcpyAsync(st.rng, opts->rng_state, (size_t)S * 32, cudaMemcpyHostToDevice, sm));
        std::vector<double> steps((size_t)S, opts->stepsize);
        CK(cudaMemcpyAsync(st.bh_step, steps.data(), (size_t)S * 8, cudaMemcpyHostToDevice, sm));
        CK(cudaStreamSynchronize(sm));  // `steps` goes out of scope
    }
    if (budget_ns > 0) {
        const int slice[4] = {ctx->fit_slice_us, 0, 0, ctx->fit_slice_us};
        CK(cudaMemcpyAsync(d_fc + FC_SLICE_US, slice, sizeof(slice), cudaMemcpyHostToDevice, sm));
        CK(cudaStreamSynchronize(sm));
    }
    const unsigned eflags = (flags | MISTI_FLAG_DEVICE_PTRS);
    const int tb = 64, gb = (S + tb - 1) / tb;
    if ((rc = ensure_batch(ctx, (size_t)B))) return rc;  // no allocation inside a round (a round may be captured)
    // cpfit post-split pass: in the JSFS kernel's lanes or in the correction chain -- decided ONCE per fit (by the size of its
    // first round), so that every point of a fit is evaluated by the same variant (they differ in the order of summation)
    const int defer = ctx->defer_post < 0 ? ((long)S * misti::nm_slots(N, false) <= kDeferPostMaxItems ? 1 : 2) : ctx->defer_post;
    // one round: propose (packs the points behind the device-side counter), evaluate, apply
    auto round_body = [&]() -> int {
        misti_fit_propose_kernel<<<gb, tb, 0, sm>>>(S, cfg, bh, st, d_par, d_bm, d_br, d_list, ctx->d_status, d_fc, look_max, B);
        CK(cudaGetLastError());
        int rc2 = eval_chunk(ctx, B, N, d_par, d_bm, -1, eflags, mixture_th, nullptr, d_llh, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, d_br, nullptr, d_fc + FC_ITEMS, defer, d_list, d_ck, d_fc + FC_SLICE_US, stable_items);
        if (rc2) return rc2;
        misti_fit_apply_kernel<<<gb, tb, 0, sm>>>(S, cfg, st, d_par, d_llh, ctx->d_status, d_fc);
        CK(cudaGetLastError());
        return 0;
    };
    // A graph = kRoundsPerGraph rounds and one copy of the counters to pinned memory; kept for the next fit as long as nothing
    // it refers to has moved.
    const int kRoundsPerGraph = ctx->nm_rounds_per_graph;
    cudaGraphExec_t exec = nullptr;
    const int64_t launches_before = ctx->launches;
    int launches_per_round = 0;
    if (ctx->nm_use_graph) {
        unsigned long long mt_bits, xa_bits, fa_bits, be_bits, ta_bits, fc_bits;
        std::memcpy(&mt_bits, &mixture_th, 8); std::memcpy(&xa_bits, &cfg.xatol, 8); std::memcpy(&fa_bits, &cfg.fatol, 8);
        std::memcpy(&be_bits, &bh.beta, 8); std::memcpy(&ta_bits, &bh.target, 8); std::memcpy(&fc_bits, &bh.factor, 8);
        const std::vector<unsigned long long> key = {ctx->generation, (unsigned long long)S, (unsigned long long)N,
                                                     (unsigned long long)cfg.lookahead, xa_bits, fa_bits,
                                                     (unsigned long long)cfg.maxiter, (unsigned long long)cfg.maxfev,
                                                     (unsigned long long)flags, mt_bits, (unsigned long long)(size_t)sm,
                                                     (unsigned long long)(long long)bh.niter, (unsigned long long)bh.interval, be_bits,
                                                     ta_bits, fc_bits, (unsigned long long)defer, (unsigned long long)kRoundsPerGraph, (unsigned long long)budget_ns, (unsigned long long)look_max};
        if (ctx->nm_graph && key == ctx->nm_graph_key) {
            exec = ctx->nm_graph;
            launches_per_round = ctx->nm_graph_launches;
        } else {
            if (ctx->nm_graph) { cudaGraphExecDestroy(ctx->nm_graph); ctx->nm_graph = nullptr; }
            if (cudaStreamBeginCapture(sm, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                int rcb = 0;
                for (int i = 0; i < kRoundsPerGraph && rcb == 0; ++i) rcb = round_body();
                if (rcb == 0 && cudaMemcpyAsync(ctx->h_nm_counts, d_fc, FC_N * sizeof(int), cudaMemcpyDeviceToHost, sm) != cudaSuccess) rcb = MISTI_E_CUDA;
                cudaGraph_t graph = nullptr;
                const cudaError_t e = cudaStreamEndCapture(sm, &graph);
                if (rcb == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    ctx->nm_graph = exec;
                    ctx->nm_graph_key = key;
                    launches_per_round = (int)((ctx->launches - launches_before) / kRoundsPerGraph) + 2;
                    ctx->nm_graph_launches = launches_per_round;
                } else {
                    exec = nullptr;
                    cudaGetLastError();  // clear the capture error: the rounds are launched directly instead
                }
                if (graph) cudaGraphDestroy(graph);
            }
            ctx->launches = launches_before;  // nothing ran during the capture
        }
    }
    volatile int* hc = ctx->h_nm_counts;
    int64_t rounds = 0, points = 0;
    if (const char* trace_path = getenv("MISTI_FIT_TRACE")) {
        // diagnostics: one round at a time, synchronised, with the device times of its kernels written to a file
        // (round, items, correction kernel ms, JSFS + stiff kernels ms, wall ms of the round)
        FILE* tf = fopen(trace_path, "a");
        for (long r = 0; r < 10000000; ++r) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            CK(cudaEventRecord(e0, sm));
            if ((rc = round_body())) return rc;
            CK(cudaEventRecord(e1, sm));
            int items = 0;
            CK(cudaMemcpyAsync(ctx->h_nm_counts, d_fc, FC_N * sizeof(int), cudaMemcpyDeviceToHost, sm));
            CK(cudaStreamSynchronize(sm));
            items = hc[FC_LAST];
            float k1 = 0, k2 = 0, all = 0;
            cudaEventElapsedTime(&k1, ctx->ev[0], ctx->ev[1]);
            cudaEventElapsedTime(&k2, ctx->ev[1], ctx->ev[2]);
            cudaEventElapsedTime(&all, e0, e1);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            if (tf) fprintf(tf, "%ld %d %.4f %.4f %.4f %d\n", r, items, k1, k2, all, hc[FC_SLICE_US]);
            if (items == 0) break;
        }
        if (tf) fclose(tf);
        exec = nullptr;
    } else
    // The host keeps two batches of rounds in flight and looks at the counters of the batch before: a last round that
    // packed nothing ends the fit (the rounds queued behind it are empty and cost microseconds).
    for (long g = 0;; ++g) {
        if (g * kRoundsPerGraph > (1L << 26))  // a fit without budgets that never converges (scipy would loop for ever as well)
            return fail(ctx, MISTI_E_ARG, "misti_fit: no end after 2^26 rounds (give maxiter / maxfev)");
        if (g >= 2) {
            CK(cudaEventSynchronize(ctx->nm_ev[(g - 2) & 3]));
            // the copy of batch g - 2 may have been overwritten by that of batch g - 1 already: the counters only grow, and
            // "nothing packed in the last round" stays true once every simplex has ended
            rounds = hc[FC_ROUNDS_USED];
            points = ((int64_t)(unsigned)hc[FC_POINTS_HI] << 32) | (unsigned)hc[FC_POINTS_LO];
            if (hc[FC_ROUND] > 0 && hc[FC_LAST] == 0) break;
        }
        if (exec) {
            CK(cudaGraphLaunch(exec, sm));
            ctx->launches += (int64_t)launches_per_round * kRoundsPerGraph;
        } else {
            for (int i = 0; i < kRoundsPerGraph; ++i) {
                if ((rc = round_body())) return rc;
                ctx->launches += 2;
            }
            CK(cudaMemcpyAsync(ctx->h_nm_counts, d_fc, FC_N * sizeof(int), cudaMemcpyDeviceToHost, sm));
        }
        CK(cudaEventRecord(ctx->nm_ev[g & 3], sm));
    }
    CK(cudaStreamSynchronize(sm));
    rounds = hc[FC_ROUNDS_USED];
    points = ((int64_t)(unsigned)hc[FC_POINTS_HI] << 32) | (unsigned)hc[FC_POINTS_LO];
    ctx->ev_valid = false;  // the timing events of the evaluation were recorded inside the rounds
    static_assert(sizeof(long long) == sizeof(int64_t), "64-bit counters");
    if (!walkers) {
        // results: best vertex and value (the simplices are sorted), counts
        CK(cudaMemcpy2DAsync(res->x, (size_t)N * 8, st.sim, (size_t)(N + 1) * N * 8, (size_t)N * 8, S, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpy2DAsync(res->fun, 8, st.fsim, (size_t)(N + 1) * 8, 8, S, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->nit, st.iters, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->nfev, st.fcalls, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->status, st.status, (size_t)S * 4, cudaMemcpyDeviceToHost, sm));
    } else {
        // results: the best minimum every walker has seen (Storage), scipy's counts
        CK(cudaMemcpyAsync(res->x, st.bh_best_x, (size_t)S * N * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->fun, st.bh_best_f, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->nit, st.bh_hop, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->nfev, st.bh_nfev, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        CK(cudaMemcpyAsync(res->status, st.bh_best_ok, (size_t)S * 4, cudaMemcpyDeviceToHost, sm));
        if (res->accepted) CK(cudaMemcpyAsync(res->accepted, st.bh_naccept, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
        if (res->failures) CK(cudaMemcpyAsync(res->failures, st.bh_fail, (size_t)S * 8, cudaMemcpyDeviceToHost, sm));
    }
    CK(cudaStreamSynchronize(sm));
    if (walkers)
        for (int s = 0; s < S; ++s) {
            res->status[s] = res->status[s] ? 0 : 1;  // 0 = the best minimum came from a converged local search
            res->nit[s] -= 1;                          // hops taken after the initial minimisation
        }
    res->rounds = rounds; res->points = points; res->graph = exec ? 1 : 0;
    return 0;
}

int misti_nelder_mead(misti_ctx* ctx, int32_t S, int32_t N, const double* x0, const int32_t* model_ids, const int32_t* row_ids,
                      uint32_t flags, double mixture_th, double xatol, double fatol, int64_t maxiter, int64_t maxfev,
                      double* x, double* fun, int64_t* nit, int64_t* nfev, int32_t* status, int64_t* info) {
    if (info) info[0] = info[1] = info[2] = 0;
    misti_fit_opts o;
    std::memset(&o, 0, sizeof(o));
    o.xatol = xatol; o.fatol = fatol; o.maxiter = maxiter; o.maxfev = maxfev; o.niter = -1;
    misti_fit_result r;
    std::memset(&r, 0, sizeof(r));
    r.x = x; r.fun = fun; r.nit = nit; r.nfev = nfev; r.status = status;
    const int rc = misti_fit(ctx, S, N, x0, model_ids, row_ids, flags, mixture_th, &o, &r);
    if (info) { info[0] = r.rounds; info[1] = r.points; info[2] = r.graph; }
    return rc;
}

int misti_score_spectra(misti_ctx* ctx, int32_t B, const double* spectra, double* llh) {
    if (!ctx) return MISTI_E_ARG;
    if (B < 0 || !spectra || !llh) return fail(ctx, MISTI_E_ARG, "misti_score_spectra: bad arguments");
    if (ctx->R < 1) return fail(ctx, MISTI_E_ARG, "misti_score_spectra: no data rows (misti_set_data)");
    if (B == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    // context-owned scratch (one block: the spectra, then the likelihoods), grown on demand
    const size_t n_sp = (size_t)B * 7, n_out = (size_t)B * ctx->R;
    int rc;
    if ((rc = ensure(ctx, &ctx->d_score, &ctx->d_score_cap, n_sp + n_out))) return rc;
    double *d_sp = ctx->d_score, *d_out = ctx->d_score + n_sp;
    CK(cudaMemcpyAsync(d_sp, spectra, n_sp * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    misti_score_kernel<<<(B + 3) / 4, 128, 0, ctx->stream>>>(B, d_sp, ctx->d_data, ctx->R, ctx->unfolded, d_out);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(llh, d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_last_kernel_ms(misti_ctx* ctx, float* out2) {
    if (!ctx || !out2) return MISTI_E_ARG;
    if (!ctx->ev_valid) return fail(ctx, MISTI_E_ARG, "misti_last_kernel_ms: no evaluation recorded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev[2]));
    CK(cudaEventElapsedTime(&out2[0], ctx->ev[0], ctx->ev[1]));
    CK(cudaEventElapsedTime(&out2[1], ctx->ev[1], ctx->ev[2]));
    return 0;
}

int64_t misti_launch_count(const misti_ctx* ctx) { return ctx ? ctx->launches : 0; }

int misti_coalescent_rates(misti_ctx* ctx, int32_t model_id, int32_t P, const double* params, double mu0, double mu1,
                           double* lh_out, double* pr_out) {
    if (!ctx) return MISTI_E_ARG;
    if (model_id < 0 || model_id >= (int)ctx->h_models.size() || P < 0 || P > MISTI_MAX_PARAMS || (P > 0 && !params) || !lh_out)
        return fail(ctx, MISTI_E_ARG, "misti_coalescent_rates: bad arguments");
    const ModelDesc& md = ctx->h_models[model_id];
    if (md.n_params > P) return fail(ctx, MISTI_E_ARG, "misti_coalescent_rates: incorrect number of parameters");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_tables(ctx))) return rc;
    const int n2 = md.splitT < md.numT ? md.splitT : md.numT;
    const size_t n_lh = 2 * (size_t)md.numT, n_pr = 6 * (size_t)(n2 + 1);
    if ((rc = ensure(ctx, &ctx->d_score, &ctx->d_score_cap, MISTI_MAX_PARAMS + n_lh + n_pr))) return rc;
    double *d_par = ctx->d_score, *d_lh = d_par + MISTI_MAX_PARAMS, *d_pr = d_lh + n_lh;
    if (P > 0) CK(cudaMemcpyAsync(d_par, params, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_pr, 0, n_pr * sizeof(double), ctx->stream));
    misti_coal_rates_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_models, model_id, ctx->d_times, ctx->d_lh, ctx->d_cls, d_par, mu0, mu1,
                                                       d_lh, d_pr);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(lh_out, d_lh, n_lh * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (pr_out) CK(cudaMemcpyAsync(pr_out, d_pr, n_pr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_generator(misti_ctx* ctx, int32_t which, double l1, double l2, double m1, double m2, double* out) {
    if (!ctx || !out || (which != 0 && which != 1)) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int n = which == 1 ? 8 : 44;
    misti_generator_kernel<<<1, 64, 0, ctx->stream>>>(which, l1, l2, m1, m2, ctx->d_small);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(out, ctx->d_small, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_pulse(misti_ctx* ctx, const double* P0, double rate, int32_t src_pop, double* P1) {
    if (!ctx || !P0 || !P1 || (src_pop != 0 && src_pop != 1)) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_small, P0, 44 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    misti_pulse_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_small, rate, src_pop, ctx->d_small + 44);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(P1, ctx->d_small + 44, 44 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_ancient_reset(misti_ctx* ctx, const double* P0, double* P1) {
    if (!ctx || !P0 || !P1) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_small, P0, 44 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    misti_ancient_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_small, ctx->d_small + 44);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(P1, ctx->d_small + 44, 44 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int misti_state_to_jaf(misti_ctx* ctx, int32_t which, int32_t* out) {
    if (!ctx || !out || (which != 0 && which != 1)) return MISTI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int n = which == 1 ? 8 : 44;
    misti_state_to_jaf_kernel<<<1, 64, 0, ctx->stream>>>(which, (int*)ctx->d_small);
    CK(cudaGetLastError());
    ctx->launches += 1;
    CK(cudaMemcpyAsync(out, ctx->d_small, (size_t)n * 7 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
