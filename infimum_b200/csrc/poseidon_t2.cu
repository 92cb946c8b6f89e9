#define INF_T 2
#include "poseidon_tu.cuh"
