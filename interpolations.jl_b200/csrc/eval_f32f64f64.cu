// eval kernels for coefficients=float, coordinates=double, weights=double (see eval.cuh)
#include "eval_combo_impl.cuh"
namespace b200itp {
template int launch_eval_combo<float, double, double>(const EvalParams &, int, int, int, cudaStream_t);
}
