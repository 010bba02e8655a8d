// The generic aggregation kernel (aggregate.cu), truncated-linear potentials.
#define MGM_GENERIC_POT 1
#include "aggregate.cu"
