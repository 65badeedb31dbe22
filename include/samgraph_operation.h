/* samgraph_operation.h — the drop-in C-ABI of the host runtime.
 *
 * Same names, argument meaning and error behaviour as the reference's
 * samgraph/common/operation.h:29-108 (definitions operation.cc:45-385); the
 * ctypes binding in samgraph/common/__init__.py:268-341 works unchanged.
 * samgraph_get_log_init_value is defined by the reference (operation.cc:267)
 * but missing from its header; samgraph_sample / samgraph_extract are declared
 * there (operation.h:95-97) but never defined — here both exist.
 * Errors: no status codes; a violated check logs and abort()s (logging.cc:69-73).
 */
#ifndef SAMGRAPH_OPERATION_H
#define SAMGRAPH_OPERATION_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void samgraph_config(const char **config_keys, const char **config_values, const size_t num_config_items); /* operation.cc:45 */
void samgraph_init(void);                 /* operation.cc:171 */
void samgraph_start(void);                /* operation.cc:179 */
void samgraph_shutdown(void);             /* operation.cc:240 */
size_t samgraph_num_epoch(void);          /* operation.cc:189 */
size_t samgraph_steps_per_epoch(void);    /* operation.cc:194 */
size_t samgraph_num_class(void);          /* operation.cc:199 */
size_t samgraph_feat_dim(void);           /* operation.cc:204 */
uint64_t samgraph_get_next_batch(void);   /* operation.cc:209 */
void samgraph_sample_once(void);          /* operation.cc:223 */
size_t samgraph_get_graph_num_src(uint64_t key, int graph_id);   /* operation.cc:225 */
size_t samgraph_get_graph_num_dst(uint64_t key, int graph_id);   /* operation.cc:230 */
size_t samgraph_get_graph_num_edge(uint64_t key, int graph_id);  /* operation.cc:235 */
void samgraph_log_step(uint64_t epoch, uint64_t step, int item, double val);      /* operation.cc:248 */
void samgraph_log_step_add(uint64_t epoch, uint64_t step, int item, double val);  /* operation.cc:254 */
void samgraph_log_epoch_add(uint64_t epoch, int item, double val);                /* operation.cc:261 */
double samgraph_get_log_init_value(int item);                                     /* operation.cc:267 */
double samgraph_get_log_step_value(uint64_t epoch, uint64_t step, int item);      /* operation.cc:272 */
double samgraph_get_log_epoch_value(uint64_t epoch, int item);                    /* operation.cc:278 */
void samgraph_report_init(void);
void samgraph_report_step(uint64_t epoch, uint64_t step);
void samgraph_report_step_average(uint64_t epoch, uint64_t step);
void samgraph_report_epoch(uint64_t epoch);
void samgraph_report_epoch_average(uint64_t epoch);
void samgraph_report_node_access(void);
void samgraph_trace_step_begin(uint64_t key, int item, uint64_t ts);
void samgraph_trace_step_end(uint64_t key, int item, uint64_t ts);
void samgraph_trace_step_begin_now(uint64_t key, int item);
void samgraph_trace_step_end_now(uint64_t key, int item);
void samgraph_dump_trace(void);
void samgraph_forward_barrier(void);
/* multi-process (arch5): data_init before fork, then per process sample_init / train_init */
void samgraph_data_init(void);                                   /* operation.cc:335 */
void samgraph_sample_init(int worker_id, const char *ctx);       /* operation.cc:343 */
void samgraph_train_init(int worker_id, const char *ctx);        /* operation.cc:350 */
void samgraph_sample(void);
void samgraph_extract(void);
void samgraph_extract_start(int count);                          /* operation.cc:357 */
void samgraph_switch_init(int worker_id, const char *ctx, double cache_percentage); /* aborts: out of scope */
size_t samgraph_num_local_step(void);                            /* operation.cc:370 */
int samgraph_wait_one_child(void);                               /* operation.cc:374 */

#ifdef __cplusplus
}
#endif
#endif /* SAMGRAPH_OPERATION_H */
