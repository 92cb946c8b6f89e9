#define INF_T 3
#include "poseidon_tu.cuh"
