// Host-callable launchers of the fast 1-D persistent kernels.  Each radix lives in its own
// translation unit so that the heavily unrolled register-FFT kernels compile in parallel.
#pragma once
#include "exb_kernels_1d.cuh"

// returns EXB_OK or an EXB_E* code (then *err points to a static message)
int exb_launch_fast1d_r16(cudaStream_t st, const exb::K1dParams<float>& p, int nscr, int max_smem, const char** err);
int exb_launch_fast1d_r8(cudaStream_t st, const exb::K1dParams<float>& p, int nscr, int max_smem, const char** err);
// true when (N, nonlinear function) has a fast instantiation
bool exb_fast1d_supported(int N, const exb::NlParams<float>& P, int order);
