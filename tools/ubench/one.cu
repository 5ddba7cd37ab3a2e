#include "field_bench.cu"
