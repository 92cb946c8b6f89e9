#define INF_T 8
#include "poseidon_tu.cuh"
