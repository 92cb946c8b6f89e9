#define INF_T 5
#include "poseidon_tu.cuh"
