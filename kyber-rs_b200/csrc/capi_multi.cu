#include "ctx.cuh"
