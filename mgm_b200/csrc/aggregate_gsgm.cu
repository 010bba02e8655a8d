// The generic aggregation kernel (aggregate.cu), SGM potentials.
#define MGM_GENERIC_POT 0
#include "aggregate.cu"
