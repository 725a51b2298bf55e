/* fcd_b200_probes.h — bring-up / micro-benchmark entry points of libfcd_b200_probes.so.
 *
 * NOT part of the product library: libfcd_b200.so does not contain these symbols.  Built on request only
 * (`python -m fcdgan_b200._build --probes`), used by scripts/gpu_probe.py, probe_halo.py, umma_bench*.py to pin the
 * tcgen05 shared-memory descriptor semantics and the UMMA issue rates the convolution kernels rely on
 * (profiles/r01_probe_halo.log, profiles/r01_umma_issue_rate.log). */
#ifndef FCD_B200_PROBES_H
#define FCD_B200_PROBES_H
#ifdef __cplusplus
extern "C" {
#endif

/* bring-up probe for the tcgen05 shared-memory descriptor semantics (scripts/gpu_probe.py); not on the product path */
int fcd_debug_umma_probe(const void* a, const void* b, float* d, int a_rows, int b_rows, int a_blocks, int b_blocks,
                         int mn_major, int n, int ksteps, int a_shift_rows, int a_base_offset, int b_shift_rows,
                         int b_base_offset, int a_sbo, int b_sbo, int a_lbo, int a_kstep, void* stream);

/* tcgen05.mma issue-rate micro-benchmark (scripts/umma_bench.py): cycles for `iters` x 4 k-steps of UMMA(s) with the given
 * shape / major-ness / descriptor geometry, operands resident in shared memory; cycles[grid].  Not on the product path. */
int fcd_debug_umma_bench(int mn_major, int n1, int n2, int a_sbo, int a_lbo, int a_kstep, int a_shift, int a2_off, int b_sbo,
                         int b_lbo, int b_kstep, int stage_stride, int stages, int b_off, int iters, int a_tmem, int grid,
                         long long* cycles, void* stream);

/* compile-time UMMA "programs" (operand-reuse experiments, scripts/umma_bench2.py; the list is in probe_tc.cu) */
int fcd_debug_umma_prog(int prog, int iters, int a_sbo, int grid, long long* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCD_B200_PROBES_H */
