/* Symbol visibility for the B200-native libcugraph_c replacement. */
#pragma once
#if defined(__GNUC__)
#define CUGRAPH_EXPORT __attribute__((visibility("default")))
#else
#define CUGRAPH_EXPORT
#endif
