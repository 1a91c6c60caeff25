// misti_jsfs.cuh -- expected joint SFS of one evaluation item, written for a cooperating GROUP of
// lanes (16 lanes of a warp on the device; a single "lane" in the test-only host build tests/hostsim).
//
// Reference path restated here: MigrationInference.JAFSpectrum / SolveDifEq / CollapsePops
// (MigrationInference.py:467-540), TwoPopulations.SetMatrix / UpdateMatrixCol / PulseMigration /
// AncientSampleP0 / StateToJAF (TwoPopulations.py:188-262, 336-377), OnePopulation.SetMatrix
// (OnePopulation.py:153-178), and the likelihood tail (MigrationInference.py:583-613).
//
// What is different from the reference (same numbers, different algorithm):
//   * intervals WITH migration: P1 = expm(M T) P0 and integralP = inv(M)(P1 - P0) = int_0^T exp(M s) P0 ds are
//     obtained together by UNIFORMISATION of the lineage chain: with q = max |M_cc| the matrix A = I + M/q is
//     non-negative, exp(M T) = sum_k Pois(k; qT) A^k, and int_0^T exp(M s) P0 ds = (1/q) sum_k P(N > k) A^k P0.
//     Every term is non-negative (no cancellation) and only sparse 44x44 mat-vecs are needed (152 off-diagonal
//     entries).  Intervals with qT > 32 are cut into equal sub-steps.
//   * RUNS of consecutive intervals WITHOUT migration: M_i = la0_i C0 + la1_i C1 with two constant, commuting,
//     diagonalisable matrices (eigenvalues 0, -1, -3, -6), so the whole run is ONE application of
//     sum_ab coef_ab G0_a G1_b (8 non-zero projector products, 188 entries; tools/gen_tables.py) with scalar
//     coefficients -- the reference's singular zero-migration case (TwoPopulations.py:240-309) in closed form.
//   * after the split all generators are multiples of one constant 8x8 matrix L8 and commute, so the whole
//     post-split contribution is sum_k cpost[k] (W8 G_k) P8 with the three spectral projectors of L8.
//   * all SCALAR work per interval (q, 1/q, Poisson start weights, series length, run coefficients) is done once per
//     item by the single thread that ran the correction chain (build_segments_item, called by kernel K1) and handed
//     to the lane group as a list of 128-byte SEGMENT RECORDS; the lane group only does vector work.
#pragma once
#include <climits>
#include <cstring>

#include "misti_model.cuh"
#include "misti_tables.h"

namespace misti {

struct EllEntry { unsigned char col, kind, cnt; };
struct PulseEntry { unsigned char row, col, a, b, mult; };

#define MISTI_DEFINE_TABLES(SPEC, PFX)                                          \
    SPEC EllEntry PFX##ell[44][MISTI_ELL_WIDTH] = MISTI_ELL_INIT;               \
    SPEC unsigned char PFX##diag[44][4] = MISTI_GEN_DIAG_INIT;                  \
    SPEC unsigned char PFX##qdiag[MISTI_QDIAG_N][4] = MISTI_QDIAG_INIT;         \
    SPEC unsigned char PFX##w44[7][44] = MISTI_W44_INIT;                        \
    SPEC unsigned char PFX##collapse[44] = MISTI_COLLAPSE_INIT;                 \
    SPEC unsigned char PFX##anc2[44] = MISTI_ANC2_INIT;                         \
    SPEC unsigned char PFX##anc11[44] = MISTI_ANC11_INIT;                       \
    SPEC PulseEntry PFX##pulse0[MISTI_PULSE0_NNZ] = MISTI_PULSE0_INIT;          \
    SPEC PulseEntry PFX##pulse1[MISTI_PULSE1_NNZ] = MISTI_PULSE1_INIT;          \
    SPEC unsigned char PFX##pulse0_rowptr[45] = MISTI_PULSE0_ROWPTR_INIT;       \
    SPEC unsigned char PFX##pulse1_rowptr[45] = MISTI_PULSE1_ROWPTR_INIT;       \
    SPEC double PFX##wg6[7][8] = MISTI_WG6_INIT;                                \
    SPEC double PFX##wg3[7][8] = MISTI_WG3_INIT;                                \
    SPEC double PFX##wg1[7][8] = MISTI_WG1_INIT;                                \
    SPEC unsigned char PFX##l16_pos[48] = MISTI_L16_POS_INIT;                   \
    SPEC unsigned char PFX##l16_row[16][3] = MISTI_L16_ROW_INIT;                \
    SPEC unsigned char PFX##l16_rem[16][3][3][2] = MISTI_L16_REM_INIT;          \
    SPEC unsigned char PFX##l16_loc[16][3][2] = MISTI_L16_LOC_INIT;             \
    SPEC unsigned char PFX##nm_rowptr[45] = MISTI_NM_ROWPTR_INIT;               \
    SPEC unsigned char PFX##nm_col[MISTI_NM_NNZ] = MISTI_NM_COL_INIT;           \
    SPEC unsigned char PFX##nm_ab[MISTI_NM_NNZ] = MISTI_NM_AB_INIT;             \
    SPEC double PFX##nm_val[MISTI_NM_NNZ] = MISTI_NM_VAL_INIT;                  \
    SPEC double PFX##r16_val[16 * MISTI_R16_LEN] = MISTI_R16_VAL_INIT;          \
    SPEC unsigned PFX##r16_meta[16 * MISTI_R16_LEN] = MISTI_R16_META_INIT;

// Under nvcc the table users are device-only functions reading __device__ copies; the test-only
// host build (g++) reads plain static copies.
#if defined(__CUDACC__)
MISTI_DEFINE_TABLES(static __device__ const, d_)
#define MISTI_TAB(name) d_##name
#define MISTI_D __device__
#else
MISTI_DEFINE_TABLES(static const, h_)
#define MISTI_TAB(name) h_##name
#define MISTI_D
#endif

// 1/k for the Poisson weight recursion p_k = p_(k-1) * lam / k (a table lookup instead of a division per term)
#define MISTI_RECIP_N 512
struct RecipTable {
    double v[MISTI_RECIP_N];
    constexpr RecipTable() : v() {
        v[0] = 0.0;
        for (int k = 1; k < MISTI_RECIP_N; ++k) v[k] = 1.0 / k;
    }
};
#if defined(__CUDACC__)
static __constant__ RecipTable c_recip = RecipTable();
#define MISTI_RECIP(k) c_recip.v[k]
#else
static const RecipTable h_recip = RecipTable();
#define MISTI_RECIP(k) h_recip.v[k]
#endif

constexpr int kYStride = 48;                    // doubles per ping-pong buffer (44 states + pad rows)
constexpr int kGroupScratch = 168;              // doubles of scratch per lane group: two ping-pong buffers + 72
constexpr int kRunOutP = 65, kRunOutI = 113;    // zero-migration run: where row results are parked (behind the staged record)
constexpr double kUnifMaxStep = 32.0;           // largest q*T handled in one uniformisation sweep
constexpr double kUnifTol = 2.2737367544323206e-13;  // 2^-42: truncation of the Poisson tail (relative to the mass).  Four thousand
                                                     // times below the 1e-9 the path is held to, and below what the other roundings of
                                                     // a sweep leave (4.5e-13 on the golden cases with 2^-50 and 2^-42 alike); 11 % fewer
                                                     // mat-vec terms than 2^-50
constexpr double kUnifMaxStiff = 256.0;         // q*T beyond this (8 sweeps) goes to the dense scaling-and-squaring step
constexpr int kUnifMaxTerms = MISTI_RECIP_N - 4;     // never reached for q*T <= 32 (about 110 terms)

// ---- segment records ----------------------------------------------------------------------------
// The two-population part of an item is a list of segments, 16 doubles (one 128-byte line) each:
//   SEG_MIG    one interval with migration, to be swept by uniformisation.  Slots 0..9: coefficient table
//              2^(c >> 2) * rate[c & 3] / q for code c = kind + 4 log2(count) (kinds: coalescence in deme 0 / 1,
//              migration out of deme 0 / 1; codes 10, 11 do not occur), 10: 1/q, 11: lam = q T / nsub, 12: 0
//              (code 12 = "no entry"), 13: Pois(0; lam), 14: P(N > 0), 15: meta.
//   SEG_RUN    a run of intervals without migration.  Slots 0..7: c_ab = sum_i e_ab(i) phi_ab(i) (integral
//              coefficients), 8..14: e_ab at the end of the run for ab = 1..7 (e_0 = 1), 15: meta.
//   SEG_STIFF  one interval with migration and q T > kUnifMaxStiff: only meta (the item is parked for the dense step).
//   SEG_INF    the infinite last interval when there is no split inside the grid: slots 0..9 table, 10: 1/q.
// meta (the 64 bits of slot 15): bits 0-2 type, bit 3 ancient-sample reset before the segment, bit 4 pulse before
// the segment, bit 5 segment lies before the sampling date, bits 8-19 first interval, bits 20-31 number of
// intervals, bits 32-47 series length K, bits 48-62 number of sub-steps.
constexpr int kRecSlots = 16;
enum { SEG_NOP = 0, SEG_MIG = 1, SEG_RUN = 2, SEG_STIFF = 3, SEG_INF = 4 };
constexpr unsigned long long kSegReset = 8, kSegPulse = 16, kSegPre = 32;

MISTI_HD inline double seg_meta_pack(unsigned long long m) { double d; memcpy(&d, &m, 8); return d; }
MISTI_HD inline unsigned long long seg_meta_bits(double d) { unsigned long long m; memcpy(&m, &d, 8); return m; }
MISTI_HD inline int seg_type(unsigned long long m) { return (int)(m & 7ull); }
MISTI_HD inline int seg_it(unsigned long long m) { return (int)((m >> 8) & 0xfffull); }
MISTI_HD inline int seg_K(unsigned long long m) { return (int)((m >> 32) & 0xffffull); }
MISTI_HD inline int seg_nsub(unsigned long long m) { return (int)((m >> 48) & 0x7fffull); }

MISTI_HD inline bool finite_nonneg(double v) { return v >= 0.0 && v <= DBL_MAX; }

// Segment list of one item (run by ONE thread, after the correction chain).  lc is addressed as
// lc[(pitch*t + g)*stride]; rec has room for min(splitT, numT) records.  Returns MISTI_OK, MISTI_NONFINITE or
// MISTI_INFINITE_COAL_TIME (no split inside the grid and no migration in the last interval, :475-476).
MISTI_D inline int build_segments_item(const ModelDesc& md, const double* times, const double* params, const double* lc,
                                       int pitch, long stride, double* rec, int* nseg_out, const unsigned* cls = nullptr) {
    const int numT = md.numT;
    const int n2 = md.splitT < numT ? md.splitT : numT;
    const bool inf_last = md.splitT >= numT;
    const int n_fin = inf_last ? n2 - 1 : n2;
    int ns = 0;
    bool open = false;
    double c[8], u0 = 1.0, u1 = 1.0;
    unsigned long long run_meta = 0;
    int run_n = 0;
    *nseg_out = 0;
    // records are written straight to their place in global memory (no staging array in the thread's stack frame)
    auto close_run = [&]() {
        double* o = rec + (long)ns * kRecSlots;
        for (int i = 0; i < 8; ++i) o[i] = c[i];
        const double u03 = u0 * u0 * u0, u13 = u1 * u1 * u1;
        o[8] = u0; o[9] = u03; o[10] = u03 * u03;
        o[11] = u1; o[12] = u13; o[13] = u13 * u13;
        o[14] = u0 * u1;
        o[15] = seg_meta_pack(run_meta | ((unsigned long long)run_n << 20));
        ++ns;
        open = false;
    };
    // coefficient table of a migration interval into slots 0..10 of record `o` (the rest zeroed); returns q
    auto table = [&](double la0, double la1, double m0, double m1, double* o) -> double {
        double q = 0.0;
        for (int i = 0; i < MISTI_QDIAG_N; ++i) {
            const double d = ((double)MISTI_TAB(qdiag)[i][0] * la0 + (double)MISTI_TAB(qdiag)[i][1] * la1) +
                             ((double)MISTI_TAB(qdiag)[i][2] * m0 + (double)MISTI_TAB(qdiag)[i][3] * m1);
            q = d > q ? d : q;
        }
        const double qinv = 1.0 / q;
        const double rq0 = la0 * qinv, rq1 = la1 * qinv, rq2 = m0 * qinv, rq3 = m1 * qinv;
        o[0] = rq0; o[1] = rq1; o[2] = rq2; o[3] = rq3;
        o[4] = 2.0 * rq0; o[5] = 2.0 * rq1; o[6] = 2.0 * rq2; o[7] = 2.0 * rq3;
        o[8] = 4.0 * rq0; o[9] = 4.0 * rq1;
        o[10] = qinv;
        o[11] = 0.0; o[12] = 0.0; o[13] = 0.0; o[14] = 0.0;
        return q;
    };
    for (int it = 0; it <= n_fin; ++it) {
        const bool last = it == n_fin;  // the infinite interval (only visited when inf_last)
        if (last && !inf_last) break;
        const int t = last ? numT - 1 : it;
        const double la0 = lc[(pitch * t) * stride], la1 = lc[(pitch * t + 1) * stride];
        double mi_t[2], pu_t[2];
        interval_rates(md, cls, params, t, mi_t, pu_t);
        const double m0 = mi_t[0], m1 = mi_t[1];
        const double T = last ? 0.0 : times[t];
        if (!(finite_nonneg(la0) && finite_nonneg(la1) && finite_nonneg(m0) && finite_nonneg(m1) && finite_nonneg(T)))
            return MISTI_NONFINITE;
        const bool reset = t == md.sampleDate && t > 0;
        const bool pulse = pu_t[0] + pu_t[1] > 0;
        const bool mig = m0 + m1 != 0.0;
        if (open && (reset || pulse || mig || last)) close_run();
        const unsigned long long flags = (reset ? kSegReset : 0) | (pulse ? kSegPulse : 0) | (t < md.sampleDate ? kSegPre : 0) |
                                         ((unsigned long long)t << 8) | (1ull << 20);
        if (last) {
            if (!mig) return MISTI_INFINITE_COAL_TIME;
            double* o = rec + (long)ns * kRecSlots;
            const double q = table(la0, la1, m0, m1, o);
            if (!(q <= DBL_MAX)) return MISTI_NONFINITE;
            o[15] = seg_meta_pack(flags | SEG_INF);
            ++ns;
        } else if (mig) {
            double* o = rec + (long)ns * kRecSlots;
            const double q = table(la0, la1, m0, m1, o);
            const double qT = q * T;
            if (!(qT <= DBL_MAX)) return MISTI_NONFINITE;
            if (qT > kUnifMaxStiff) {
                // rates of 1e5 and more per unit of interval length only come out of a run-away correction; such an
                // interval is not swept (it would take millions of terms): the item is parked for the dense step
                for (int i = 0; i < 11; ++i) o[i] = 0.0;
                o[15] = seg_meta_pack(flags | SEG_STIFF);
            } else {
                const int nsub = qT > kUnifMaxStep ? (int)ceil(qT / kUnifMaxStep) : 1;
                const double lam = nsub == 1 ? qT : qT / nsub;
                const double p0 = exp(-lam), t0 = -expm1(-lam);  // Poisson P(N = 0) and P(N > 0)
                // series length: the first k whose Poisson tail beyond the term is below kUnifTol
                double p = p0, r = lam;
                int k = 0;
                do {
                    ++k;
                    p *= r;
                    r = lam * MISTI_RECIP(k + 1);
                } while (!(r < 1.0 && p < kUnifTol * (1.0 - r)) && k < kUnifMaxTerms);
                if (k >= kUnifMaxTerms) return MISTI_NONFINITE;
                o[11] = lam; o[13] = p0; o[14] = t0;
                o[15] = seg_meta_pack(flags | SEG_MIG | ((unsigned long long)k << 32) | ((unsigned long long)nsub << 48));
            }
            ++ns;
        } else {
            if (!open) {
                open = true;
                u0 = 1.0; u1 = 1.0;
                for (int i = 0; i < 8; ++i) c[i] = 0.0;
                run_meta = (flags & ~(0xfffull << 20)) | SEG_RUN;
                run_n = 0;
            }
            ++run_n;
            // e_ab = exp(-a X0 - b X1) at the start of the interval; phi_ab = (1 - exp(-(a la0 + b la1) T)) / (a la0 + b la1);
            // 1 - v^3 = (1 - v)(1 + v + v^2) and 1 - v0 v1 = (1 - v0) + v0 (1 - v1) keep every term positive
            const double z0 = la0 * T, z1 = la1 * T;
            // v = exp(-z) as 1 - w: the absolute error (1e-16) is what matters for a survival factor
            const double w0 = -expm1(-z0), w1 = -expm1(-z1), v0 = 1.0 - w0, v1 = 1.0 - w1;
            const double v03 = v0 * v0 * v0, v13 = v1 * v1 * v1;
            const double w03 = w0 * (1.0 + v0 + v0 * v0), w13 = w1 * (1.0 + v1 + v1 * v1);
            const double w06 = w03 * (1.0 + v03), w16 = w13 * (1.0 + v13);
            const double u03 = u0 * u0 * u0, u13 = u1 * u1 * u1;
            const double i0 = 1.0 / la0, i1 = 1.0 / la1, i01 = 1.0 / (la0 + la1);
            c[0] += T;
            c[1] += u0 * (la0 > 0.0 ? w0 * i0 : T);
            c[2] += u03 * (la0 > 0.0 ? w03 * (i0 * (1.0 / 3.0)) : T);
            c[3] += (u03 * u03) * (la0 > 0.0 ? w06 * (i0 * (1.0 / 6.0)) : T);
            c[4] += u1 * (la1 > 0.0 ? w1 * i1 : T);
            c[5] += u13 * (la1 > 0.0 ? w13 * (i1 * (1.0 / 3.0)) : T);
            c[6] += (u13 * u13) * (la1 > 0.0 ? w16 * (i1 * (1.0 / 6.0)) : T);
            c[7] += (u0 * u1) * (la0 + la1 > 0.0 ? (w0 + v0 * w1) * i01 : T);
            u0 *= v0; u1 *= v1;
        }
    }
    if (open) close_run();
    *nseg_out = ns;
    return MISTI_OK;
}

// ---- lane groups ------------------------------------------------------------------------------
// A group = the lanes that share one item.  Several groups may share a warp; they then run in LOCK
// STEP: every loop bound and branch in jsfs_item is made warp-uniform with wmax / wmin / any, and a
// group that has nothing to do in a step executes it with neutral weights (an exact no-op), so that the
// result of an item never depends on which item shares its warp.
#if !defined(__CUDACC__)
static const double kNeutralRec[kRecSlots] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 0};

struct SingleLane {  // test-only host build: one lane owns all 44 rows, every off-diagonal entry is "remote"
    static constexpr int LANES = 1;
    static constexpr int RPL = 44;    // rows per lane
    static constexpr int NLOC = 0;    // entries served from the lane's own registers, per row
    static constexpr int RWMAX = 4;   // entries read from the shared vector, per row (upper bound)
    struct Rec {
        const double* p;
        MISTI_HD double get(int i) const { return p[i]; }
    };
    MISTI_D static int rw(int) { return 4; }
    static constexpr int RUNLEN = MISTI_NM_NNZ;  // entries of the run table per lane
    MISTI_D int row_of(int s) const { return s; }
    MISTI_D static int pos_of(int row) { return row; }
    MISTI_D void rem_of(int s, int e, int* col, unsigned* code) const {
        const EllEntry en = MISTI_TAB(ell)[s][e];
        *col = en.col;
        *code = en.cnt == 0 ? 12u : (unsigned)en.kind + (en.cnt == 1 ? 0u : (en.cnt == 2 ? 4u : 8u));
    }
    MISTI_D unsigned loc_of(int, int) const { return 12u; }
    MISTI_HD int lane() const { return 0; }
    MISTI_HD void sync() const {}
    MISTI_HD double sum(double v) const { return v; }
    MISTI_HD double excl_prod(double) const { return 1.0; }
    MISTI_HD int wmax(int v) const { return v; }
    MISTI_HD int wmin(int v) const { return v; }
    MISTI_HD bool any(bool v) const { return v; }
    MISTI_HD bool any_in_group(bool v) const { return v; }
    // record j of the item (all-zero = SEG_NOP when the group has none)
    MISTI_HD Rec rec_load(const double* p, bool have) const { return Rec{have ? p : zero_rec()}; }
    // the record itself, or the neutral sweep record (no events, P(N = 0) = 1)
    MISTI_HD Rec rec_select(const Rec& r, bool is) const { return Rec{is ? r.p : kNeutralRec}; }
    MISTI_HD void rec_store(const Rec& r, double* dst) const {  // 16 slots and the constant 1 behind them
        for (int i = 0; i < kRecSlots; ++i) dst[i] = r.p[i];
        dst[kRecSlots] = 1.0;
    }
    MISTI_HD static const double* zero_rec() {
        static const double z[kRecSlots] = {0};
        return z;
    }
};

#else
struct HalfWarpLanes {  // two items per warp; each lane of a 16-lane half owns three states (slots a, b, c) chosen so
                        // that many off-diagonal entries have their column in the same lane (tools/gen_tables.py)
    static constexpr int LANES = 16;
    static constexpr int RPL = 3;
    static constexpr int NLOC = 2;   // per row: one coefficient for each of the lane's other two rows (registers)
    static constexpr int RWMAX = 3;  // per row: at most 3 / 2 / 3 entries read from shared memory (slots a / b / c)
    struct Rec {  // lane l of the group holds slot l; get() is one shuffle
        double v;
        __device__ double get(int i) const { return __shfl_sync(0xffffffffu, v, i, 16); }
    };
    __device__ static int rw(int s) { return s == 1 ? 2 : 3; }
    static constexpr int RUNLEN = MISTI_R16_LEN;
    __device__ int row_of(int s) const { return d_l16_row[threadIdx.x & 15][s]; }
    __device__ static int pos_of(int row) { return d_l16_pos[row]; }  // shared-memory word of a state: 16 * slot + lane
    __device__ void rem_of(int s, int e, int* col, unsigned* code) const {
        *col = d_l16_rem[threadIdx.x & 15][s][e][0];
        *code = d_l16_rem[threadIdx.x & 15][s][e][1];
    }
    __device__ unsigned loc_of(int s, int j) const { return d_l16_loc[threadIdx.x & 15][s][j]; }
    __device__ int lane() const { return threadIdx.x & 15; }
    __device__ void sync() const { __syncwarp(); }
    __device__ double sum(double v) const {  // within the half
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ double excl_prod(double v) const {  // product of v over the lower lanes of the half (1 for lane 0)
        const int lane = threadIdx.x & 15;
        for (int o = 1; o < 16; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, o, 16);
            if (lane >= o) v *= t;
        }
        const double e = __shfl_up_sync(0xffffffffu, v, 1, 16);
        return lane == 0 ? 1.0 : e;
    }
    __device__ int wmax(int v) const {  // over the whole warp (both items); v is uniform within a half
        const int o = __shfl_xor_sync(0xffffffffu, v, 16);
        return v > o ? v : o;
    }
    __device__ int wmin(int v) const {
        const int o = __shfl_xor_sync(0xffffffffu, v, 16);
        return v < o ? v : o;
    }
    __device__ bool any(bool v) const { return __any_sync(0xffffffffu, v); }
    __device__ bool any_in_group(bool v) const { return (__ballot_sync(0xffffffffu, v) & (0xffffu << (threadIdx.x & 16u))) != 0; }
    __device__ Rec rec_load(const double* p, bool have) const { return Rec{have ? p[threadIdx.x & 15] : 0.0}; }
    __device__ Rec rec_select(const Rec& r, bool is) const { return Rec{is ? r.v : ((threadIdx.x & 15) == 13 ? 1.0 : 0.0)}; }
    __device__ void rec_store(const Rec& r, double* dst) const {
        dst[threadIdx.x & 15] = r.v;
        if ((threadIdx.x & 15) == 15) dst[kRecSlots] = 1.0;
    }
};
#endif

// Continuation record of an item whose next segment is a stiff interval (rates ~1e5..1e8 after a run-away
// correction): the state at the START of that interval (before its ancient reset / pulse).  misti_stiff_kernel
// advances it over the stiff segment(s) with a dense scaling-and-squaring step; the JSFS kernel then resumes from it.
struct Cont {
    int seg;
    int nterms;
    int pad_[2];
    double P[48];
    double Ia[48];
    double Ib[48];
};

// Post-split coefficients (run by ONE thread; lc addressed like in correct_lambdas_item):
//   cpost[k] = sum_{i>=splitT} exp(-a_k x_i) (1 - exp(-a_k lam_i T_i)) / (a_k lam_i),  x_i = sum_{j<i} lam_j T_j,
// with the last interval infinite (MigrationInference.py:530-540: P1 = 0 there), a = (6, 3, 1).
MISTI_HD inline void post_split_coeffs(const ModelDesc& md, const double* times, const double* lc, int pitch, long stride,
                                       double* cpost) {
    double c6 = 0, c3 = 0, c1 = 0;
    double e1 = 1.0;  // exp(-x)
    for (int t = md.splitT; t < md.numT; ++t) {
        const double lam = lc[(pitch * t) * stride];
        const double e3 = e1 * e1 * e1, e6 = e3 * e3;
        if (t < md.numT - 1) {
            const double z = lam * times[t];
            const double u = exp(-z), w1 = -expm1(-z);      // w1 = 1 - u
            const double w3 = w1 * (1.0 + u + u * u);         // 1 - u^3
            const double w6 = w3 * (1.0 + u * u * u);         // 1 - u^6
            const double il = 1.0 / lam;
            c1 += e1 * w1 * il;
            c3 += e3 * w3 * (il * (1.0 / 3.0));
            c6 += e6 * w6 * (il * (1.0 / 6.0));
            e1 *= u;
        } else {
            const double il = 1.0 / lam;
            c1 += e1 * il;
            c3 += e3 * (il * (1.0 / 3.0));
            c6 += e6 * (il * (1.0 / 6.0));
        }
    }
    cpost[0] = c6; cpost[1] = c3; cpost[2] = c1;
}

// The same coefficients in cpfit mode, straight from ed = exp(nc1 - nc0) of the correction chain (see
// post_split_cpfit_item in misti_model.cuh: the post-split intervals are independent of each other given ed), by ALL lanes
// of a group together: lane l takes the `per` consecutive intervals of its slice of the model's table (post_split_table:
// tab[(k per + j) LANES + l]), accumulates its partial sums relative to the start of the slice, the survival factors
// exp(-x) at the slice starts come from one exclusive scan over the lanes, and the last lane adds the infinite interval
// (lh_last = PSMC rates of that interval).  Every lane returns the three sums.  Called by all lanes of the warp (`active`
// = false for a group without an item: it runs along with neutral values).
template <class G>
MISTI_D inline void post_split_cpfit_group(const G& g, bool active, const double* tab, int per, const double* lh_last, double ed,
                                           double* cpost) {
    const int lane = g.lane();
    const int per_own = active ? per : 0;
    const int per_max = g.wmax(per_own);
    const double wn = 1.0 / (1.0 + ed);
    double c6 = 0, c3 = 0, c1 = 0, e1 = 1.0;  // relative to the start of this lane's slice
    // the loads of step j + 1 are in flight while step j is computed (the table may sit in global memory)
    const long ks = (long)per * G::LANES;
    const double* q = tab + lane;
    double nE0 = 1.0, nE1 = 1.0, nQ0 = 0.0, nQ1 = 0.0, nT = 0.0;
    if (0 < per_own) { nE0 = q[0]; nE1 = q[ks]; nQ0 = q[2 * ks]; nQ1 = q[3 * ks]; nT = q[4 * ks]; }
    for (int j = 0; j < per_max; ++j) {
        const double E0 = nE0, E1 = nE1, Q0 = nQ0, Q1 = nQ1, T = nT;
        nE0 = 1.0; nE1 = 1.0; nQ0 = 0.0; nQ1 = 0.0; nT = 0.0;
        if (j + 1 < per_own) {
            q += G::LANES;
            nE0 = q[0]; nE1 = q[ks]; nQ0 = q[2 * ks]; nQ1 = q[3 * ks]; nT = q[4 * ks];
        }
        const double u = (E0 + ed * E1) * wn;   // exp(-lam T), the fitted non-coalescence probability
        const double q1 = (Q0 + ed * Q1) * wn;  // 1 - u, free of cancellation
        const double z = -log(u);
        const double il = z > 0 ? T / z : 0.0;  // 1 / lam
        const double e3 = e1 * e1 * e1;
        const double q3 = q1 * (1.0 + u + u * u), q6 = q3 * (1.0 + u * u * u);  // 1 - u^3, 1 - u^6
        c1 += z > 0 ? e1 * q1 * il : e1 * T;
        c3 += z > 0 ? e3 * q3 * (il * (1.0 / 3.0)) : e3 * T;
        c6 += z > 0 ? (e3 * e3) * q6 * (il * (1.0 / 6.0)) : (e3 * e3) * T;
        e1 *= u;
    }
    const double f1 = g.excl_prod(e1), f3 = f1 * f1 * f1;
    c1 *= f1; c3 *= f3; c6 *= f3 * f3;
    if (active && lane == G::LANES - 1) {  // the infinite last interval
        const double lam = (1.0 + ed) / (1.0 / lh_last[0] + ed / lh_last[1]);
        const double il = 1.0 / lam, x1 = f1 * e1, x3 = x1 * x1 * x1;
        c1 += x1 * il; c3 += x3 * (il * (1.0 / 3.0)); c6 += (x3 * x3) * (il * (1.0 / 6.0));
    }
    cpost[0] = g.sum(c6); cpost[1] = g.sum(c3); cpost[2] = g.sum(c1);
}

// The zero-migration run table as the lane groups use it.  Any lane may compute any row (inputs and outputs go through
// shared memory), so the 188 entries G_ab[row][col] are dealt into LANES lists of equal length, row by row; the lanes
// walk their lists in lock step, entry k of lane l at index k * LANES + l (consecutive words: no bank conflicts).
// meta = shared-memory word of `col` | slot of c_ab << 8 | slot of e_ab << 16 in the staged record (e_0 = 1 sits behind
// the record, slot 16) | "last entry of its row" << 24 | shared-memory word of the row << 25.  For 16 lanes the table is
// generated (tools/gen_tables.py: balanced lists, order chosen for few gather conflicts); a single lane takes the rows
// in their natural order.
template <class G>
struct RunTable {
    double val[G::RUNLEN * G::LANES];
    unsigned meta[G::RUNLEN * G::LANES];
    MISTI_D void fill(int tid, int nthreads) {
        if (G::LANES == 16) {
            for (int i = tid; i < G::RUNLEN * G::LANES; i += nthreads) { val[i] = MISTI_TAB(r16_val)[i]; meta[i] = MISTI_TAB(r16_meta)[i]; }
        } else {
            for (int r = tid; r < 44; r += nthreads)
                for (int e = MISTI_TAB(nm_rowptr)[r]; e < MISTI_TAB(nm_rowptr)[r + 1]; ++e) {
                    const unsigned ab = MISTI_TAB(nm_ab)[e], last = e + 1 == MISTI_TAB(nm_rowptr)[r + 1];
                    val[e] = MISTI_TAB(nm_val)[e];
                    meta[e] = (unsigned)MISTI_TAB(nm_col)[e] | (ab << 8) | ((ab == 0 ? (unsigned)kRecSlots : 7u + ab) << 16) | (last << 24) |
                              ((unsigned)r << 25);
                }
        }
    }
};

// What a lane knows about its rows, independent of the item: built once per thread.
template <class G>
struct LaneCtx {
    static constexpr int RPL = G::RPL, NLOC = G::NLOC, RW = G::RWMAX;
    int row[RPL];        // the states owned by this lane (44..47 = pad rows of empty slots)
    unsigned rc[RPL];    // coefficient-table codes (kind + 4 log2(count), 12 = none): remote entry e at bit 4e, local entry j at bit 16+4j
    unsigned rk[RPL];    // diagonal multiplicity of rate kind k at bit 3k; StateToJAF count of category c at bit 12+2c;
                         // collapse block at bit 26; ancient-reset masks at bits 29, 30; valid at bit 31
    const double* yp[RPL][RW];  // where this lane reads y[col] of each remote entry (buffer 0; buffer 1 is +kYStride)
    double* wp[RPL];            // where it writes y[row]
    double* scratch;            // the group's scratch area (kGroupScratch doubles)
    const RunTable<G>* runtab;

    MISTI_D void init(const G& g, double* ysm, const RunTable<G>* rt) {
        scratch = ysm;
        runtab = rt;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            row[s] = g.row_of(s);
            const bool valid = row[s] < 44;
            wp[s] = ysm + G::pos_of(row[s]);
            rc[s] = 0; rk[s] = 0;
#pragma unroll
            for (int e = 0; e < RW; ++e) {
                int col = row[s];
                unsigned cde = 12u;
                if (valid && e < G::rw(s)) g.rem_of(s, e, &col, &cde);
                rc[s] |= cde << (4 * e);
                yp[s][e] = ysm + G::pos_of(col);
            }
#pragma unroll
            for (int j = 0; j < NLOC; ++j) rc[s] |= (valid ? g.loc_of(s, j) : 12u) << (16 + 4 * j);
            if (valid) {
                const int r = row[s];
                for (int k = 0; k < 4; ++k) rk[s] |= (unsigned)MISTI_TAB(diag)[r][k] << (3 * k);
                for (int c = 0; c < 7; ++c) rk[s] |= (unsigned)MISTI_TAB(w44)[c][r] << (12 + 2 * c);
                rk[s] |= (unsigned)MISTI_TAB(collapse)[r] << 26;
                rk[s] |= (unsigned)MISTI_TAB(anc2)[r] << 29;
                rk[s] |= (unsigned)MISTI_TAB(anc11)[r] << 30;
                rk[s] |= 1u << 31;
            }
        }
    }
};

// Expected JSFS of one item.  ALL lanes of the warp call this together (each group with its own item;
// `active` = false for a group without work); `rec` / `nseg` = the item's segment records (build_segments_item);
// L.scratch is private to the group.  On return lane c (and, with 16 lanes, c + 8) of the group holds the
// UNNORMALISED spectrum entry c in *jafs_c (MigrationInference.JAFSpectrum's return value; lanes 7 and 15 hold 0;
// a single lane gets all of them in jafs_c[0..6]) and every lane the number of mat-vecs in *terms.
// `cont` (nullable): where to park the item when it meets a stiff segment (return value MISTI_STIFF = "pending");
// `resume`: start from the record instead of from the sampling configuration.
template <class G>
MISTI_D inline int jsfs_item(const G& g, const LaneCtx<G>& L, const ModelDesc& md, bool active, const double* params,
                             const double* rec, int nseg, const double* cpost, double* jafs_c, int* terms,
                             Cont* cont = nullptr, bool resume = false) {
    constexpr int RPL = G::RPL, NLOC = G::NLOC, RW = G::RWMAX;
    typedef typename G::Rec Rec;
    const int lane = g.lane();
    double* const ysm = L.scratch;
    double P[RPL];            // state probabilities at the start of the current segment (rows owned by this lane)
    double Ia[RPL], Ib[RPL];  // occupancy integrals summed over the intervals before / from the sampling date
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        P[s] = L.row[s] == 2 ? 1.0 : 0.0;  // both genome-1 lineages in deme 0, genome-2 in deme 1 (:469-471)
        Ia[s] = 0.0; Ib[s] = 0.0;
    }
    int nterms = 0;
    int status = MISTI_OK;
    int seg0 = 0;
    if (resume && active && cont) {
        seg0 = cont->seg;
        nterms = cont->nterms;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const bool valid = (L.rk[s] >> 31) != 0;
            P[s] = valid ? cont->P[L.row[s]] : 0.0;
            Ia[s] = valid ? cont->Ia[L.row[s]] : 0.0;
            Ib[s] = valid ? cont->Ib[L.row[s]] : 0.0;
        }
    }
    bool pending = false;
    const int numT = md.numT;
    const int nown = active ? nseg : 0;

    // AncientSampleP0 (TwoPopulations.py:246-262): mass of the states with both genome-1 singletons in deme 0 -> state 2,
    // with a (2,0) lineage in deme 0 -> state 11, the rest is dropped
    auto ancient_reset = [&](bool do_reset) {
        double a2 = 0.0, a11 = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            if ((L.rk[s] >> 29) & 1u) a2 += P[s];
            if ((L.rk[s] >> 30) & 1u) a11 += P[s];
        }
        a2 = g.sum(a2); a11 = g.sum(a11);
        if (do_reset) {
#pragma unroll
            for (int s = 0; s < RPL; ++s) P[s] = L.row[s] == 2 ? a2 : (L.row[s] == 11 ? a11 : 0.0);
        }
    };
    // the reset, then PulseMigration (:361-377), before interval `it`
    auto reset_and_pulse = [&](int it, bool do_reset, bool do_pulse) {
        if (g.any(do_reset)) ancient_reset(do_reset);
        if (g.any(do_pulse)) {  // a group without a pulse here applies the map with rate 0 = the identity
            double pr = 0.0;
            int src = 0;
            if (do_pulse) {
                const double pu0 = pulse_rate(md, params, it, 0), pu1 = pulse_rate(md, params, it, 1);
                pr = pu0 + pu1;
                src = pu0 > 0 ? 0 : 1;
            }
            const double om = 1.0 - pr;
            const PulseEntry* ent = src == 0 ? MISTI_TAB(pulse0) : MISTI_TAB(pulse1);
            const unsigned char* rp = src == 0 ? MISTI_TAB(pulse0_rowptr) : MISTI_TAB(pulse1_rowptr);
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s) L.wp[s][0] = P[s];
            g.sync();
            double pw_om[5], pw_r[5];
            pw_om[0] = 1.0; pw_r[0] = 1.0;
            for (int k = 1; k < 5; ++k) { pw_om[k] = pw_om[k - 1] * om; pw_r[k] = pw_r[k - 1] * pr; }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                double acc = 0.0;
                if (L.rk[s] >> 31) {
                    const int r = L.row[s];
                    for (int e = rp[r]; e < rp[r + 1]; ++e) {
                        const PulseEntry pe = ent[e];
                        double w = (double)pe.mult;
                        for (int k = 0; k < 5; ++k) {
                            if (k == pe.a) w *= pw_om[k];
                            if (k == pe.b) w *= pw_r[k];
                        }
                        acc += w * ysm[G::pos_of(pe.col)];
                    }
                }
                P[s] = acc;
            }
            g.sync();
        }
    };

    // generator in uniformised form, A = I + M/q, from the coefficient table of a record
    double adiag[RPL], coef[RPL][RW], cloc[RPL][NLOC > 0 ? NLOC : 1];
    auto load_generator = [&](const Rec& sw) {
        const double rq0 = sw.get(0), rq1 = sw.get(1), rq2 = sw.get(2), rq3 = sw.get(3);
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const unsigned dg = L.rk[s];
            const double d = ((double)(dg & 7u) * rq0 + (double)((dg >> 3) & 7u) * rq1) +
                             ((double)((dg >> 6) & 7u) * rq2 + (double)((dg >> 9) & 7u) * rq3);
            adiag[s] = d < 1.0 ? 1.0 - d : 0.0;
#pragma unroll
            for (int e = 0; e < RW; ++e) coef[s][e] = sw.get((L.rc[s] >> (4 * e)) & 15u);
#pragma unroll
            for (int j = 0; j < NLOC; ++j) cloc[s][j] = sw.get((L.rc[s] >> (16 + 4 * j)) & 15u);
        }
    };
    // y_new[row] = (A y)[row] for the rows of this lane: diagonal and same-lane entries from registers (yk), the rest
    // from the shared copy of y at offset RO
    auto matvec = [&](const double* yk, const int RO, double* acc) {
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            double a = adiag[s] * yk[s];
            if (NLOC == 2) {
                a = fma(cloc[s][0], yk[(s + 1) % RPL], a);
                a = fma(cloc[s][NLOC - 1], yk[(s + 2) % RPL], a);
            }
#pragma unroll
            for (int e = 0; e < RW; ++e)
                if (e < G::rw(s)) a = fma(coef[s][e], L.yp[s][e][RO], a);
            acc[s] = a;
        }
    };

    // one interval with migration: uniformisation sweep(s)
    auto sweep = [&](const Rec& rv, bool is, unsigned long long meta) {
        const Rec sw = g.rec_select(rv, is);
        load_generator(sw);
        const double qinv = sw.get(10), lam_own = sw.get(11), p0_own = sw.get(13), t0_own = sw.get(14);
        const int K_own = is ? seg_K(meta) : 0, nsub_own = is ? seg_nsub(meta) : 0;
        const int nsub = g.wmax(nsub_own);
        double Iint[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) Iint[s] = 0.0;
        for (int sub = 0; sub < nsub; ++sub) {
            const bool live = sub < nsub_own;
            double lam = live ? lam_own : 0.0;
            const double p0 = live ? p0_own : 1.0, t0 = live ? t0_own : 0.0;
            const int Ks = live ? K_own : 0;
            const int Kmin = g.wmin(live ? Ks : INT_MAX);  // some group is live in every sub-step
            const int Kmax = g.wmax(Ks);
            double yk[RPL], P1[RPL];
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                yk[s] = P[s]; P1[s] = p0 * P[s];
                L.wp[s][0] = P[s];
            }
            double p = p0;     // Pois(k; lam)
            double tail = t0;  // P(N > k)
            double r = lam;    // lam / (k + 1): ratio of consecutive Poisson weights
            int k = 0;
            // one term: I += P(N > k-1) y_(k-1);  y_k <- A y_(k-1) (read buffer RO, write buffer WO);  P1 += Pois(k) y_k
            auto term = [&](const int RO, const int WO) {
                g.sync();
                ++k;
                p *= r;
                r = lam * MISTI_RECIP(k + 1);
                double acc[RPL];
                matvec(yk, RO, acc);
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    Iint[s] = fma(tail, yk[s], Iint[s]);
                    yk[s] = acc[s];
                    L.wp[s][WO] = acc[s];
                    P1[s] = fma(p, acc[s], P1[s]);
                }
                tail -= p;
            };
            // the terms every group needs: the two halves of the ping-pong buffer get compile-time offsets
            while (k + 2 <= Kmin) {
                term(0, kYStride);
                term(kYStride, 0);
            }
            int cur = 0;
            if (k < Kmin) { term(0, kYStride); cur = kYStride; }
            // the rest: a group whose own series has ended goes on with zero weights (exact no-op)
            while (k < Kmax) {
                if (k >= Ks) { lam = 0.0; r = 0.0; tail = 0.0; }
                term(cur, kYStride - cur);
                cur = kYStride - cur;
            }
            if (live) nterms += Ks;
#pragma unroll
            for (int s = 0; s < RPL; ++s) P[s] = P1[s];
        }
        // integralP of this interval joins the running sums; categories 2..6 are muted before the sampling
        // date (:501-506), hence the two accumulators
        const bool pre = (meta & kSegPre) != 0;
        const double qa = pre ? qinv : 0.0, qb = pre ? 0.0 : qinv;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            Ia[s] = fma(Iint[s], qa, Ia[s]);
            Ib[s] = fma(Iint[s], qb, Ib[s]);
        }
    };

    // a run of intervals without migration: P <- sum_ab e_ab G_ab P, integral += sum_ab c_ab G_ab P
    auto runop = [&](const Rec& rv, bool is, unsigned long long meta) {
        g.sync();
#pragma unroll
        for (int s = 0; s < RPL; ++s) L.wp[s][0] = P[s];
        g.rec_store(rv, ysm + kYStride);
        g.sync();
        if (is) {  // this lane's share of the table; a row's sums are parked when its last entry has been added
            const double* cc = ysm + kYStride;
            double pe = 0.0, ir = 0.0;
#pragma unroll 4
            for (int k = 0; k < G::RUNLEN; ++k) {
                const unsigned m = L.runtab->meta[k * G::LANES + lane];
                const double t = L.runtab->val[k * G::LANES + lane] * ysm[m & 255u];
                ir = fma(cc[(m >> 8) & 255u], t, ir);
                pe = fma(cc[(m >> 16) & 255u], t, pe);
                if ((m >> 24) & 1u) {
                    ysm[kRunOutP + (m >> 25)] = pe;
                    ysm[kRunOutI + (m >> 25)] = ir;
                    pe = 0.0; ir = 0.0;
                }
            }
        }
        g.sync();
        if (is) {
            const bool pre = (meta & kSegPre) != 0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (L.rk[s] >> 31) {
                    P[s] = L.wp[s][kRunOutP];
                    if (pre) Ia[s] += L.wp[s][kRunOutI];
                    else Ib[s] += L.wp[s][kRunOutI];
                }
            nterms += 1;
        }
    };

    // no split inside the grid: the last two-population interval is infinite (MigrationInference.py:475-476,
    // 535-538): P1 = 0, integralP = -inv(M) P0 = (1/q) sum_k A^k P0 (finite only with migration)
    auto infsum = [&](const Rec& rv, bool is, unsigned long long meta) {
        const Rec sw = g.rec_select(rv, is);
        load_generator(sw);
        const double qinv = sw.get(10);
        double yk[RPL], Iint[RPL];
        g.sync();
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            yk[s] = is ? P[s] : 0.0;
            Iint[s] = yk[s];
            L.wp[s][0] = yk[s];
        }
        double nprev = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) nprev += yk[s];
        nprev = g.sum(nprev);
        double itot = nprev;
        bool done = !(nprev > 0.0);
        int k = 0, cur = 0;
        while (g.any(!done)) {
            g.sync();
            const int ro = kYStride * cur, wo = kYStride * (cur ^ 1);
            double nk = 0.0, acc[RPL];
            matvec(yk, ro, acc);
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                if (done) acc[s] = 0.0;  // this group's series has ended: keep its sums as they are
                yk[s] = acc[s];
                L.wp[s][wo] = acc[s];
                Iint[s] += acc[s];
                nk += acc[s];
            }
            cur ^= 1;
            ++k;
            nk = g.sum(nk);
            itot += nk;
            const double rho = nprev > 0.0 ? nk / nprev : 0.0;  // contraction of the remaining mass
            nprev = nk;
            if (!done && (!(nk > 0.0) || (rho < 1.0 && nk * rho < kUnifTol * itot * (1.0 - rho)))) {
                done = true;
                if (is) nterms += k;
            }
            if (k > 2000000) {
                if (!done) status = MISTI_NONFINITE;
                break;
            }
        }
        if (is) {
            const bool pre = (meta & kSegPre) != 0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                if (pre) Ia[s] = fma(Iint[s], qinv, Ia[s]);
                else Ib[s] = fma(Iint[s], qinv, Ib[s]);
                P[s] = 0.0;
            }
        }
    };

    const int n_loop = g.wmax(nown > seg0 ? nown - seg0 : 0);
    Rec nxt = g.rec_load(rec + (long)seg0 * kRecSlots, seg0 < nown);
    for (int j = 0; j < n_loop; ++j) {
        const int sg = seg0 + j;
        const bool have = !pending && sg < nown;  // a group past its own last segment idles
        const Rec rv = nxt;
        nxt = g.rec_load(rec + (long)(sg + 1) * kRecSlots, sg + 1 < nown);  // in flight while this segment is processed
        const double mslot = rv.get(15);
        const unsigned long long meta = have ? seg_meta_bits(mslot) : 0ull;
        int type = seg_type(meta);
        const int it = seg_it(meta);
        if (type == SEG_STIFF) {
            if (cont) {
                if (lane == 0) { cont->seg = sg; cont->nterms = nterms; }
#pragma unroll
                for (int s = 0; s < RPL; ++s) { cont->P[L.row[s]] = P[s]; cont->Ia[L.row[s]] = Ia[s]; cont->Ib[L.row[s]] = Ib[s]; }
                pending = true;
            } else {
                status = MISTI_STIFF;
            }
            type = SEG_NOP;
        }
        const bool do_reset = type != SEG_NOP && (meta & kSegReset) != 0, do_pulse = type != SEG_NOP && (meta & kSegPulse) != 0;
        if (g.any(do_reset || do_pulse)) reset_and_pulse(it, do_reset, do_pulse);
        if (g.any(type == SEG_MIG)) sweep(rv, type == SEG_MIG, meta);
        if (g.any(type == SEG_RUN)) runop(rv, type == SEG_RUN, meta);
        if (g.any(type == SEG_INF)) infsum(rv, type == SEG_INF, meta);
    }

    // JAFS = StateToJAF . (sum of the interval integrals) (:501-506) + the one-population tail after the split: with
    // P8 = CollapsePops(P) (:518-528) that is sum_b V[c][b] P8[b], V = c6 WG6 + c3 WG3 + c1 WG1, folded here into
    // per-row weights V[c][block(row)] so that one reduction per category does both parts.
    const bool post = active && md.splitT < numT;
    const bool any_post = g.any(post);
    if (any_post) {
        const bool do_reset = post && md.splitT == md.sampleDate && md.splitT > 0;  // the reset precedes the collapse (:480-494)
        if (g.any(do_reset)) ancient_reset(do_reset);
        const double c6 = post ? cpost[0] : 0.0, c3 = post ? cpost[1] : 0.0, c1 = post ? cpost[2] : 0.0;
        g.sync();
        for (int idx = lane; idx < 56; idx += G::LANES) {
            const int c = idx >> 3, b = idx & 7;
            ysm[idx] = (c6 * MISTI_TAB(wg6)[c][b] + c3 * MISTI_TAB(wg3)[c][b]) + c1 * MISTI_TAB(wg1)[c][b];
        }
        g.sync();
    }
    double jl[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) jl[c] = 0.0;
#pragma unroll
    for (int s = 0; s < RPL; ++s)
        if (L.rk[s] >> 31) {
            const double iab = Ia[s] + Ib[s];
            const int blk = (int)((L.rk[s] >> 26) & 7u);
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const double w = (double)((L.rk[s] >> (12 + 2 * c)) & 3u);
                jl[c] = fma(w, c < 2 ? iab : Ib[s], jl[c]);
                if (any_post) jl[c] = fma(ysm[c * 8 + blk], P[s], jl[c]);
            }
        }
    // transposed reduction through the scratch area: lane l parks its 7 partial sums, lane c adds up category c
    g.sync();
#pragma unroll
    for (int c = 0; c < 7; ++c) ysm[c * G::LANES + lane] = jl[c];
    g.sync();
    if (G::LANES == 1) {
        for (int c = 0; c < 7; ++c) jafs_c[c] = ysm[c];
    } else {
        const int c = lane & 7;
        double v = 0.0;
        if (c < 7) {
#pragma unroll
            for (int j = 0; j < G::LANES; ++j) v += ysm[c * G::LANES + ((j + c) & (G::LANES - 1))];  // rotated: no bank conflict
        }
        jafs_c[0] = v;
    }
    *terms = nterms;
    return pending ? MISTI_STIFF : status;
}

// Group-cooperative tail: normalise the spectrum, take the logs (lane c handles category c), and leave
// raw[7] / jn[7] / logj[7] in the group's scratch area at offsets kTailRaw / kTailJn / kTailLog for every lane to read.
// Returns false if a required log is not finite.  (MigrationInference.py:583-613)
constexpr int kTailRaw = 96, kTailJn = 104, kTailLog = 112;

// Normalised spectrum -> log terms used by the composite likelihood (MigrationInference.py:583-613).
// Folded: bins (0+6), (1+5), (2+4), 3; the data vector is folded the same way by the host, so
// logj[4..6] = 0 there.  Returns false if a required log is not finite.
MISTI_HD inline bool jafs_normalise_logs(const double* raw, bool unfolded, double* jn, double* logj) {
    double tot = 0.0;
    for (int c = 0; c < 7; ++c) tot += raw[c];
    for (int c = 0; c < 7; ++c) jn[c] = raw[c] / tot;
    bool ok = true;
    if (unfolded) {
        for (int c = 0; c < 7; ++c) logj[c] = log(jn[c]);
    } else {
        logj[0] = log(jn[0] + jn[6]);
        logj[1] = log(jn[1] + jn[5]);
        logj[2] = log(jn[2] + jn[4]);
        logj[3] = log(jn[3]);
        logj[4] = logj[5] = logj[6] = 0.0;
    }
    for (int c = 0; c < 7; ++c)
        if (!(fabs(logj[c]) <= DBL_MAX)) ok = false;
    return ok;
}

// llh for one data row: const + sum_i d_i log p_i, accumulated in the reference's order (:600-609).
MISTI_HD inline double score_row(const double* drow /* 7 counts (folded by the host if needed) + const */, const double* logj) {
    double llh = drow[7];
    for (int c = 0; c < 7; ++c) llh += drow[c] * logj[c];
    return llh;
}

// Group-cooperative version of the two functions above: lane c normalises and takes the log of category c
// (jafs_c as returned by jsfs_item); raw[7], jn[7], logj[7] are left in the scratch area for every lane to read, and
// *jn_c is this lane's normalised entry.
template <class G>
MISTI_D inline bool jafs_finish(const G& g, double* ysm, const double* jafs_c, bool unfolded, double* jn_c) {
    if (G::LANES == 1) {
        double jn[7], logj[7];
        const bool ok = jafs_normalise_logs(jafs_c, unfolded, jn, logj);
        for (int c = 0; c < 7; ++c) { ysm[kTailRaw + c] = jafs_c[c]; ysm[kTailJn + c] = jn[c]; ysm[kTailLog + c] = logj[c]; }
        *jn_c = jn[0];
        return ok;
    }
    const int lane = g.lane(), c = lane & 7;
    const bool mine = lane < 7;
    const double raw = jafs_c[0];
    const double tot = g.sum(mine ? raw : 0.0);
    const double jn = raw / tot;
    g.sync();  // the partial sums of jsfs_item have been read
    if (mine) { ysm[kTailRaw + c] = raw; ysm[kTailJn + c] = jn; }
    g.sync();
    double lj = 0.0;
    if (mine) {
        if (unfolded || c == 3) lj = log(jn);
        else if (c < 3) lj = log(jn + ysm[kTailJn + 6 - c]);
        ysm[kTailLog + c] = lj;
    }
    const bool bad = mine && !(fabs(lj) <= DBL_MAX);
    g.sync();
    *jn_c = jn;
    return !g.any_in_group(bad);
}

}  // namespace misti
