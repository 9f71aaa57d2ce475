// Second translation unit of the trace kernels: the backward half of trace.cu (its launch code, ab200_trace_bwd and the
// blocker-gradient reduction), compiled in parallel with the forward half - trace.cu as ONE unit took 3 min 40 s of the
// in-tree build, the two halves take about half of that side by side.  Device code and templates are shared by inclusion;
// each half instantiates only the kernels it launches.
#define AB200_TU_BWD 1
#include "trace.cu"
