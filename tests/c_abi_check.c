/* Compiled by tests/test_host_cpu.py with the host C compiler: the public header must be plain C, and the layout of
 * the argument structs must be what octfusion_b200/_lib.py declares to ctypes. */
#include <stdio.h>
#include <stddef.h>
#include "octfusion_b200.h"

int main(void) {
  printf("sizeof_gemm_args %zu\n", sizeof(of_gemm_args));
  printf("off_tap_tab %zu\n", offsetof(of_gemm_args, tap_tab));
  printf("off_w %zu\n", offsetof(of_gemm_args, w));
  printf("off_out %zu\n", offsetof(of_gemm_args, out));
  printf("off_M %zu\n", offsetof(of_gemm_args, M));
  printf("off_a_multi %zu\n", offsetof(of_gemm_args, a_multi));
  printf("off_nt_block %zu\n", offsetof(of_gemm_args, nt_block));
  printf("off_reverse %zu\n", offsetof(of_gemm_args, reverse));
  printf("off_stat_out %zu\n", offsetof(of_gemm_args, stat_out));
  printf("off_stat_rows_per_sample %zu\n", offsetof(of_gemm_args, stat_rows_per_sample));
  printf("sizeof_octree_levels %zu\n", sizeof(of_octree_levels));
  printf("off_nnum %zu\n", offsetof(of_octree_levels, nnum));
  printf("off_full_depth %zu\n", offsetof(of_octree_levels, full_depth));
  return 0;
}
