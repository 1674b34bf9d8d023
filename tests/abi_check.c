/* abi_check.c — compiled with plain gcc against the headers under include/ by tests/test_abi_cpu.py: prints sizeof and the
 * offset of every field of every struct the C-ABI exchanges, as the C compiler lays them out.  The test
 * compares the output with the ctypes mirrors in c4a0_b200/_lib.py, so a reordered or resized field in a
 * header (or a drifted mirror) fails on a CPU-only box.  Also proves the headers are valid C99. */
#include <stddef.h>
#include <stdio.h>

#include "c4a0_engine.h"
#include "c4a0_net.h"

#define S(T) printf("%s %zu\n", #T, sizeof(T))
#define F(T, f) printf("%s.%s %zu %zu\n", #T, #f, offsetof(T, f), sizeof(((T *)0)->f))

int main(void) {
  S(c4a0_config);
  F(c4a0_config, n_slots);
  F(c4a0_config, max_requests);
  F(c4a0_config, n_mcts_iterations);
  F(c4a0_config, c_exploration);
  F(c4a0_config, c_ply_penalty);
  F(c4a0_config, plane_dtype);
  F(c4a0_config, max_inline_sims);
  F(c4a0_config, device);
  F(c4a0_config, plane_stride);
  F(c4a0_config, flags);
  F(c4a0_config, arena_blocks);
  F(c4a0_config, eval_cache_entries);
  F(c4a0_config, spec_rows);
  F(c4a0_config, dirichlet_alpha);
  F(c4a0_config, dirichlet_epsilon);
  S(c4a0_progress);
  F(c4a0_progress, n_requests);
  F(c4a0_progress, n_started);
  F(c4a0_progress, n_finished);
  F(c4a0_progress, n_running);
  F(c4a0_progress, n_movers);
  F(c4a0_progress, n_rows);
  F(c4a0_progress, error);
  S(c4a0_stats);
  F(c4a0_stats, sims);
  F(c4a0_stats, nn_evals);
  F(c4a0_stats, leaf_requests);
  F(c4a0_stats, terminal_leaf_sims);
  F(c4a0_stats, skipped_root_sims);
  F(c4a0_stats, moves);
  F(c4a0_stats, samples);
  F(c4a0_stats, select_depth_sum);
  F(c4a0_stats, expansions);
  F(c4a0_stats, steps);
  F(c4a0_stats, compacted_blocks);
  F(c4a0_stats, compactions);
  F(c4a0_stats, cache_hits);
  F(c4a0_stats, cache_inserts);
  F(c4a0_stats, spec_rows);
  S(c4a0_nn_graph);
  F(c4a0_nn_graph, rows);
  F(c4a0_nn_graph, graph_exec);
  S(c4a0_run_report);
  F(c4a0_run_report, ticks);
  F(c4a0_run_report, nn_launches);
  F(c4a0_run_report, nn_rows_launched);
  F(c4a0_run_report, nn_relaunches);
  F(c4a0_run_report, tail_launches);
  F(c4a0_run_report, device_ms);
  F(c4a0_run_report, wall_ms);
  F(c4a0_run_report, host_wait_ms);
  F(c4a0_run_report, host_launch_ms);
  F(c4a0_run_report, kernel_samples);
  F(c4a0_run_report, reserved);
  F(c4a0_run_report, k_step_ms_sum);
  F(c4a0_run_report, nn_ms_sum);
  F(c4a0_run_report, bucket_launches);
  S(c4a0_slot_info);
  F(c4a0_slot_info, state);
  F(c4a0_slot_info, request);
  F(c4a0_slot_info, n_moves);
  F(c4a0_slot_info, root_visits);
  F(c4a0_slot_info, root_mask);
  F(c4a0_slot_info, root_value);
  F(c4a0_slot_info, root_q_sum_penalty);
  F(c4a0_slot_info, root_q_sum_no_penalty);
  F(c4a0_slot_info, n_blocks);
  F(c4a0_slot_info, nn_row);
  S(c4a0_net_layer);
  F(c4a0_net_layer, weight_dev);
  F(c4a0_net_layer, bias_dev);
  F(c4a0_net_layer, n_pad);
  F(c4a0_net_layer, k_pad);
  F(c4a0_net_layer, in_buffer);
  F(c4a0_net_layer, in_col0);
  F(c4a0_net_layer, out_buffer);
  F(c4a0_net_layer, out_col0);
  F(c4a0_net_layer, dep);
  F(c4a0_net_layer, kind);
  S(c4a0_net_spec);
  F(c4a0_net_spec, device);
  F(c4a0_net_spec, max_rows);
  F(c4a0_net_spec, n_buffers);
  F(c4a0_net_spec, buffer_cols);
  F(c4a0_net_spec, planes_buffer);
  F(c4a0_net_spec, planes_col0);
  F(c4a0_net_spec, n_layers);
  F(c4a0_net_spec, layers);
  printf("C4A0_ABI_VERSION %d\n", C4A0_ABI_VERSION);
  printf("C4A0_MAX_SAMPLES %d\n", C4A0_MAX_SAMPLES);
  printf("C4A0_NET_MAX_LAYERS %d\n", C4A0_NET_MAX_LAYERS);
  printf("C4A0_NET_PAD_N %d\n", C4A0_NET_PAD_N);
  return 0;
}
