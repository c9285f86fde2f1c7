// kernels.cuh -- launchers shared between the translation units of libqgsb.
#pragma once
#include "common.cuh"

// Largest ndim served by the thread-per-member generic kernel; above it a warp works on one member.
#define QGSB_G1_MAX_NDIM 64

namespace qgsb {

void launch_aos_to_soa(const double *d_in, double *d_out, long N, int n, long ld);
void launch_soa_to_aos(const double *d_in, double *d_out, long N, int n, long ld);
// (R, rows, ld) member-minor records -> (N, rows, R) API layout
void launch_rec_to_api(const double *d_in, double *d_out, long N, long rows, long R, long ld, int flip);

// sums and sums of squares per variable of R tiled records (R, n, ld) over the first N members -> (R, n) each
void launch_record_moments(const double *d_rec, long R, long N, int n, long ld, double *d_sum, double *d_sumsq);

inline long round_up(long x, long m) { return (x + m - 1) / m * m; }

// Butcher tableau checked and reduced on the host
struct Tableau {
    int s = 0;
    std::vector<double> a, b;   // a (s, s) row-major with only j < i kept
    bool chain = false;         // a_ij != 0 only for j == i - 1
    std::vector<double> alpha;  // chain form: alpha[i] = a[i][i-1]
};
Tableau make_tableau(int s, const double *a, const double *b);

// advance an (n, ld) member-minor state; d_rec == nullptr -> no recording
void rk_advance(const qgsb_tensor *t, double *d_y, long ld, long N, long n_steps, const double *d_dt,
                const Tableau &tab, long write_steps, long R, double *d_rec);

}  // namespace qgsb
