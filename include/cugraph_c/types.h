/*
 * Scalar vocabulary of the C ABI.  The enumerator names, their order and their numeric values are the ABI of the
 * reference's cpp/include/cugraph_c/types.h:14-32 and must not change; everything else in this file is ours.
 *
 * Which of the ids this library accepts where:
 *   vertex / edge ids   INT32 or INT64 (graph create, BFS sources, PageRank vertex lists); internal ids are always 32-bit
 *   edge weights        FLOAT32 or FLOAT64; an unweighted graph computes and reports FLOAT32
 *   array views         any id below NTYPES is a legal element type for create / copy; algorithms check what they need
 */
#pragma once
#include <cugraph_c/export.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* C has no bool across an ABI: an int-sized enum, as the reference does. */
typedef enum bool_ {
  FALSE = 0,
  TRUE  = 1
} bool_t;

/* raw storage unit of type-erased host arrays */
typedef int8_t byte_t;

typedef enum data_type_id_ {
  INT8    = 0,  /* 1 byte  */
  INT16   = 1,  /* 2 bytes */
  INT32   = 2,  /* 4 bytes: the vertex / edge id type of the _v32_e32 configuration */
  INT64   = 3,  /* 8 bytes: accepted for external ids; renumbering maps them to 32-bit internal ids */
  UINT8   = 4,
  UINT16  = 5,
  UINT32  = 6,
  UINT64  = 7,
  FLOAT32 = 8,  /* weights, PageRank scores, SSSP distances */
  FLOAT64 = 9,
  SIZE_T  = 10, /* 8 bytes on every platform this builds for */
  BOOL    = 11, /* 1 byte */
  NTYPES  = 12  /* number of ids, not a type */
} cugraph_data_type_id_t;

#ifdef __cplusplus
}
#endif
