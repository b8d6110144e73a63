// TEST HARNESS ONLY: compiles sfa_numpy_b200/csrc/sfa_math.cuh for the host so
// the FP64 recurrences can be checked against the golden vectors without a GPU.
// Never loaded by the product path (sfa_numpy_b200/ has no reference to it).
#include <vector>
#include "../../sfa_numpy_b200/csrc/sfa_math.cuh"

using namespace sfa;

struct RowOut {
    c128* row;
    void operator()(int c, c128 v) const { row[c] = v; }
};

extern "C" {

void hm_sph_radial(int N, long F, const double* k, double r, double rs, int setup, int scale, c128* out) {
    std::vector<double> a(N + 4), b(N + 4);
    for (long f = 0; f < F; ++f) {
        dvec jx{a.data(), 1}, js{b.data(), 1};
        sph_radial_row(N, k[f], r, rs, setup, scale, jx, js, RowOut{out + f * (N + 1)});
    }
}

void hm_circ_radial(int N, long F, const double* k, double r, double rs, int setup, int scale, c128* out) {
    std::vector<double> a(N + 4), b(N + 4);
    for (long f = 0; f < F; ++f) {
        dvec jx{a.data(), 1}, js{b.data(), 1};
        circ_radial_row(N, k[f], r, rs, setup, scale, jx, js, RowOut{out + f * (2 * N + 1)});
    }
}

void hm_regularize(long n, const c128* dn, double a0, int method, c128* dn_out, double* hn) {
    const double alpha = (method == REG_TIKH) ? tikh_alpha(a0) : 0.0;
    for (long i = 0; i < n; ++i) regularize_one(dn[i], a0, method, alpha, dn_out + i, hn + i);
}

void hm_regularize_inverse(long n, const c128* bn, double a0, int method, c128* dn_out, double* hn) {
    const double alpha = (method == REG_TIKH) ? tikh_alpha(a0) : 0.0;
    for (long i = 0; i < n; ++i) regularize_inverse(bn[i], a0, method, alpha, dn_out + i, hn + i);
}

void hm_sht(int N, long Q, const double* azi, const double* colat, int colat_scalar, const double* w, c128* Y) {
    const long M = (long)(N + 1) * (N + 1);
    for (long q = 0; q < Q; ++q)
        sh_row_scalar(N, azi[q], colat[colat_scalar ? 0 : q], w ? w[q] : 1.0, Y + q * M);
}

void hm_bessel(int nmax, double x, double* j, double* J, double* Y01) {
    dvec a{j, 1}, b{J, 1};
    sph_jn_array(nmax, x, a);
    cyl_jn_array(nmax, x, b, Y01, Y01 + 1);
}
}
