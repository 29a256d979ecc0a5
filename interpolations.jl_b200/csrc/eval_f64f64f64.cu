// eval kernels for coefficients=double, coordinates=double, weights=double (see eval.cuh)
#include "eval_combo_impl.cuh"
namespace b200itp {
template int launch_eval_combo<double, double, double>(const EvalParams &, int, int, int, cudaStream_t);
}
