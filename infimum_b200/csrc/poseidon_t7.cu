#define INF_T 7
#include "poseidon_tu.cuh"
