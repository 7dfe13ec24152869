/* TEST INFRASTRUCTURE: the helper functions the reference's C-API tests expect from their test_utils library
 * (prototypes: cpp/tests/c_api/c_test_utils.h:34-75), written here against this repository's C ABI so that
 * cpp/tests/c_api/{pagerank,bfs,sssp}_test.c compile UNMODIFIED from where they lie and run against libcugraph_c.
 * Behaviour restated from cpp/tests/c_api/test_utils.cpp: run_sg_test :243-266 (prints RUNNING/passed/FAILED, returns the
 * test's code), nearlyEqual :42-57 (|a-b| <= max(|a|,|b|) * eps), create_test_graph :61-150 (device copies of the host
 * COO, cugraph_graph_create_with_times_sg, is_multigraph FALSE, no drop/symmetrize flags, expensive check FALSE). */
#include <cugraph_c/algorithms.h>
#include <cugraph_c/graph.h>

#include <math.h>
#include <stdio.h>
#include <time.h>

int run_sg_test(int (*test)(void), const char* test_name)
{
  printf("RUNNING: %s...", test_name);
  fflush(stdout);
  time_t t0 = time(NULL);
  int rc    = test();
  printf("done (%f seconds). - %s\n", difftime(time(NULL), t0), rc == 0 ? "passed" : "FAILED");
  fflush(stdout);
  return rc;
}

int run_sg_test_new(int (*test)(const cugraph_resource_handle_t*), const char* test_name, const cugraph_resource_handle_t* handle)
{
  printf("RUNNING: %s...", test_name);
  fflush(stdout);
  int rc = test(handle);
  printf("done. - %s\n", rc == 0 ? "passed" : "FAILED");
  fflush(stdout);
  return rc;
}

int nearlyEqual(float a, float b, float epsilon)
{
  const float m = fabsf(a) < fabsf(b) ? fabsf(b) : fabsf(a);
  return fabsf(a - b) <= m * epsilon;
}

int nearlyEqualDouble(double a, double b, double epsilon)
{
  const double m = fabs(a) < fabs(b) ? fabs(b) : fabs(a);
  return fabs(a - b) <= m * epsilon;
}

static int to_device(const cugraph_resource_handle_t* h, const void* host, size_t n, cugraph_data_type_id_t t,
                     cugraph_type_erased_device_array_t** arr, cugraph_type_erased_device_array_view_t** view,
                     cugraph_error_t** err)
{
  if (cugraph_type_erased_device_array_create(h, n, t, arr, err) != CUGRAPH_SUCCESS) return 1;
  *view = cugraph_type_erased_device_array_view(*arr);
  return cugraph_type_erased_device_array_view_copy_from_host(h, *view, (const byte_t*)host, err) != CUGRAPH_SUCCESS;
}

static int make_graph(const cugraph_resource_handle_t* h, int32_t* src, int32_t* dst, void* wgt, cugraph_data_type_id_t wt,
                      size_t n, bool_t store_transposed, bool_t renumber, bool_t is_symmetric, cugraph_graph_t** g,
                      cugraph_error_t** err)
{
  cugraph_graph_properties_t props;
  props.is_symmetric  = is_symmetric;
  props.is_multigraph = FALSE;
  cugraph_type_erased_device_array_t *a_src = NULL, *a_dst = NULL, *a_wgt = NULL;
  cugraph_type_erased_device_array_view_t *v_src = NULL, *v_dst = NULL, *v_wgt = NULL;
  int bad = to_device(h, src, n, INT32, &a_src, &v_src, err) || to_device(h, dst, n, INT32, &a_dst, &v_dst, err) ||
            to_device(h, wgt, n, wt, &a_wgt, &v_wgt, err);
  if (!bad) {
    cugraph_error_code_t rc = cugraph_graph_create_with_times_sg(h, &props, NULL, v_src, v_dst, v_wgt, NULL, NULL, NULL, NULL,
                                                                 store_transposed, renumber, FALSE, FALSE, FALSE, FALSE, g, err);
    if (rc != CUGRAPH_SUCCESS) {
      printf("ASSERTION FAILED: graph creation failed: %s\n", cugraph_error_message(*err));
      bad = 1;
    }
  }
  if (v_wgt) cugraph_type_erased_device_array_view_free(v_wgt);
  if (v_dst) cugraph_type_erased_device_array_view_free(v_dst);
  if (v_src) cugraph_type_erased_device_array_view_free(v_src);
  if (a_wgt) cugraph_type_erased_device_array_free(a_wgt);
  if (a_dst) cugraph_type_erased_device_array_free(a_dst);
  if (a_src) cugraph_type_erased_device_array_free(a_src);
  return bad;
}

int create_test_graph(const cugraph_resource_handle_t* h, int32_t* src, int32_t* dst, float* wgt, size_t n,
                      bool_t store_transposed, bool_t renumber, bool_t is_symmetric, cugraph_graph_t** g, cugraph_error_t** err)
{
  return make_graph(h, src, dst, wgt, FLOAT32, n, store_transposed, renumber, is_symmetric, g, err);
}

int create_test_graph_double(const cugraph_resource_handle_t* h, int32_t* src, int32_t* dst, double* wgt, size_t n,
                             bool_t store_transposed, bool_t renumber, bool_t is_symmetric, cugraph_graph_t** g,
                             cugraph_error_t** err)
{
  return make_graph(h, src, dst, wgt, FLOAT64, n, store_transposed, renumber, is_symmetric, g, err);
}
