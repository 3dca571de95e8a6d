// Host-callable launchers of the fast 2-D / 3-D pass kernels (own translation units).
#pragma once
#include "exb_kernels_nd.cuh"

// true when (D, N, nonlinear function) has fast instantiations (f32 only)
bool exb_fastnd_supported(int D, int N, const exb::NlParams<float>& P);
// p.mode selects COL_PLAIN / COL_INV_PRO / COL_FWD_EPI / COL_FWD_NL; dir = -1 forward, +1 inverse
int exb_fastnd_col(cudaStream_t st, const exb::ColParams<float>& p, int dir, long long grid_units, const char** err);
// p.mode selects ROW_NL / ROW_R2C / ROW_C2R
int exb_fastnd_row(cudaStream_t st, const exb::RowParams<float>& p, const char** err);
// Per-pass twiddle table of the register FFT for line length N (Fft8Tw<N>), arranged on the host: `size` entries
// (0 when N has no fast kernels); exb_fastnd_tw_arrange fills out[0 .. size) from the N roots of unity.  The plan
// stores it right behind the roots (p.tw + N).
int exb_fastnd_tw_size(int N);
void exb_fastnd_tw_arrange(int N, const exb::cpx<float>* roots, exb::cpx<float>* out);
