/* Umbrella header, as the reference's cpp/include/cugraph_c/algorithms.h (hot-path subset). */
#pragma once
#include <cugraph_c/centrality_algorithms.h>
#include <cugraph_c/graph_functions.h>
#include <cugraph_c/labeling_algorithms.h>
#include <cugraph_c/traversal_algorithms.h>
