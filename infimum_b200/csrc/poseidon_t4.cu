#define INF_T 4
#include "poseidon_tu.cuh"
