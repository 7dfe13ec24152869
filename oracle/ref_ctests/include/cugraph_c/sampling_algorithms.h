/* TEST INFRASTRUCTURE (oracle/ref_ctests): the reference's c_test_utils.h includes this header for two opaque types that
 * appear in one prototype (validate_sample_result).  Sampling is outside the PageRank/BFS/SSSP path, so the
 * product's include/ tree does not carry it; this stand-in only lets the reference's own C tests compile unmodified. */
#pragma once
#include <cugraph_c/types.h>
#include <stdint.h>
typedef struct { int32_t align_; } cugraph_sample_result_t;
typedef struct { int32_t align_; } cugraph_sampling_options_t;
