// tests/hostsim -- TEST-ONLY host build of the __host__ __device__ numerics in
// misti_b200/csrc/misti_math.cuh and misti_model.cuh, so the scalar correction chain (K1's body)
// can be unit-tested against the oracle in a container without a GPU.  It is NOT part of the
// product: libmisti_b200.so has no CPU path and nothing under misti_b200/ loads this library.
#include <cstring>
#include <vector>
#include "../../misti_b200/csrc/misti_model.cuh"
#include "../../misti_b200/csrc/misti_jsfs.cuh"

extern "C" {

void hs_expm3(const double* A, double* E) { misti::mat3_expm(A, E); }

int hs_inv3(const double* A, double* Ainv) { return misti::mat3_inv(A, Ainv) ? 1 : 0; }

// bands: [n][5] = pop(0/1), start, end, value, opt index (-1 fixed); pulses: [n][4] = pop, time, value, opt index
int hs_correct_lambdas(int numT, int splitT, int sampleDate, const double* times, const double* lh, int n_bands,
                       const double* bands, int n_pulses, const double* pulses, int n_params, const double* params,
                       unsigned flags, double mixtureTH, double* lc, double* Pr, int* nfev) {
    misti::ModelDesc md;
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = splitT; md.sampleDate = sampleDate;
    md.n_bands = n_bands; md.n_pulses = n_pulses; md.n_params = n_params;
    for (int b = 0; b < n_bands; ++b) {
        md.band_pop[b] = (int)bands[5 * b]; md.band_start[b] = (int)bands[5 * b + 1]; md.band_end[b] = (int)bands[5 * b + 2];
        md.band_val[b] = bands[5 * b + 3]; md.band_opt[b] = (int)bands[5 * b + 4];
    }
    for (int b = 0; b < n_pulses; ++b) {
        md.pulse_pop[b] = (int)pulses[4 * b]; md.pulse_time[b] = (int)pulses[4 * b + 1];
        md.pulse_val[b] = pulses[4 * b + 2]; md.pulse_opt[b] = (int)pulses[4 * b + 3];
    }
    // the grid constants and interval classes the library precomputes on the host (misti_add_grid / misti_add_model)
    std::vector<double> gaux((size_t)numT * misti::kGridAux);
    for (int t = 0; t < numT; ++t) misti::grid_aux_row(lh + 2 * t, t < numT - 1 ? times[t] : 0.0, &gaux[(size_t)t * misti::kGridAux]);
    std::vector<unsigned> cls(numT);
    for (int t = 0; t < numT; ++t) cls[t] = misti::interval_class(md, t);
    return misti::correct_lambdas_item(md, times, lh, params, flags, mixtureTH, lc, 2, 1, Pr, nfev, gaux.data(), nullptr, nullptr,
                                       cls.data());
}

static void fill_model(misti::ModelDesc& md, int numT, int splitT, int sampleDate, int n_bands, const double* bands,
                       int n_pulses, const double* pulses, int n_params) {
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = splitT; md.sampleDate = sampleDate;
    md.n_bands = n_bands; md.n_pulses = n_pulses; md.n_params = n_params;
    for (int b = 0; b < n_bands; ++b) {
        md.band_pop[b] = (int)bands[5 * b]; md.band_start[b] = (int)bands[5 * b + 1]; md.band_end[b] = (int)bands[5 * b + 2];
        md.band_val[b] = bands[5 * b + 3]; md.band_opt[b] = (int)bands[5 * b + 4];
    }
    for (int b = 0; b < n_pulses; ++b) {
        md.pulse_pop[b] = (int)pulses[4 * b]; md.pulse_time[b] = (int)pulses[4 * b + 1];
        md.pulse_val[b] = pulses[4 * b + 2]; md.pulse_opt[b] = (int)pulses[4 * b + 3];
    }
}

static int hs_types_buf[256];
static int* hs_last_types = hs_types_buf;
static int hs_last_nseg = 0;

// segment types of the last hs_jsfs call (1 = swept interval, 2 = closed-form run, 3 = stiff, 4 = infinite)
int hs_segment_types(int* out, int cap) {
    for (int i = 0; i < hs_last_nseg && i < cap; ++i) out[i] = hs_types_buf[i];
    return hs_last_nseg;
}

// expected JSFS (unnormalised raw[7], normalised jn[7]) and llh for one data row, given lc[numT][2]
int hs_jsfs(int numT, int splitT, int sampleDate, const double* times, int n_bands, const double* bands, int n_pulses,
            const double* pulses, int n_params, const double* params, const double* lc, int unfolded,
            const double* drow /* 7 + const */, double* raw, double* jn, double* llh, int* terms) {
    misti::ModelDesc md;
    fill_model(md, numT, splitT, sampleDate, n_bands, bands, n_pulses, pulses, n_params);
    double cpost[3], ysm[misti::kGroupScratch], logj[7];
    misti::post_split_coeffs(md, times, lc, 2, 1, cpost);
    std::vector<double> rec((size_t)(numT + 1) * misti::kRecSlots);
    int nseg = 0;
    std::vector<unsigned> cls(numT);
    for (int t = 0; t < numT; ++t) cls[t] = misti::interval_class(md, t);
    int st = misti::build_segments_item(md, times, params, lc, 2, 1, rec.data(), &nseg, cls.data());
    if (st != MISTI_OK) return st;
    if (hs_last_types) {
        for (int i = 0; i < nseg && i < 256; ++i) hs_last_types[i] = misti::seg_type(misti::seg_meta_bits(rec[i * misti::kRecSlots + 15]));
        hs_last_nseg = nseg;
    }
    misti::SingleLane g;
    static misti::RunTable<misti::SingleLane> runtab;
    runtab.fill(0, 1);
    misti::LaneCtx<misti::SingleLane> L;
    L.init(g, ysm, &runtab);
    st = misti::jsfs_item<misti::SingleLane>(g, L, md, true, params, rec.data(), nseg, cpost, raw, terms);
    if (st != MISTI_OK) return st;
    double jn0;
    if (!misti::jafs_finish(g, ysm, raw, unfolded != 0, &jn0)) return MISTI_NONFINITE;
    for (int c = 0; c < 7; ++c) { jn[c] = ysm[misti::kTailJn + c]; logj[c] = ysm[misti::kTailLog + c]; }
    *llh = misti::score_row(drow, logj);
    return MISTI_OK;
}

}  // extern "C"
