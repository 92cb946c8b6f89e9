#define INF_T 6
#include "poseidon_tu.cuh"
