// eval kernels for coefficients=double, coordinates=float, weights=double (see eval.cuh)
#include "eval_combo_impl.cuh"
namespace b200itp {
template int launch_eval_combo<double, float, double>(const EvalParams &, int, int, int, cudaStream_t);
}
