// eval kernels for coefficients=double, coordinates=float, weights=float (see eval.cuh)
#include "eval_combo_impl.cuh"
namespace b200itp {
template int launch_eval_combo<double, float, float>(const EvalParams &, int, int, int, cudaStream_t);
}
