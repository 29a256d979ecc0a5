// Instantiates eval_kernel for one (coefficient, coordinate, weight) type combination; included by
// the eval_<combo>.cu translation units so that they compile in parallel.
#pragma once
#include "eval.cuh"

namespace b200itp {

template <typename TC, typename TX, typename TW>
int launch_eval_combo(const EvalParams &P, int deg, int mode, int blocks, cudaStream_t st) {
#define B200_CASE(NN, DD, MM)                                         \
  if (P.ndims == NN && deg == DD && mode == MM) {                     \
    eval_kernel<NN, DD, TC, TX, TW, MM><<<blocks, 256, 0, st>>>(P);   \
    return 0;                                                         \
  }
#define B200_CASES_N(NN) \
  B200_CASE(NN, 0, 0) B200_CASE(NN, 1, 0) B200_CASE(NN, 2, 0) B200_CASE(NN, 3, 0) \
  B200_CASE(NN, 0, 1) B200_CASE(NN, 1, 1) B200_CASE(NN, 2, 1) B200_CASE(NN, 3, 1) \
  B200_CASE(NN, 0, 2) B200_CASE(NN, 1, 2) B200_CASE(NN, 2, 2) B200_CASE(NN, 3, 2)
  B200_CASES_N(1)
  B200_CASES_N(2)
  B200_CASES_N(3)
  B200_CASES_N(4)
#undef B200_CASES_N
#undef B200_CASE
  return B200ITP_E_UNSUPPORTED;
}

}  // namespace b200itp
