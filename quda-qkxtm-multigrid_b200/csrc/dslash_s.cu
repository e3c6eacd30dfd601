#include "dslash_launch.cuh"
namespace qb { QB_INSTANTIATE_DSLASH(StoreS) }
