// engine_internal.h -- what the other translation units of libgkquad_b200.so may touch of a handle
// (gkq_handle_s itself is private to engine.cu).  Not part of the ABI: hidden visibility.
#pragma once
#include "../../include/gkquad_b200.h"

struct gkq_comm_s;  // comm.cu

#define GKQ_HIDDEN __attribute__((visibility("hidden")))
GKQ_HIDDEN int gkq_internal_fail(gkq_handle_t h, int code, const char* msg);  // h == NULL: the creation error
GKQ_HIDDEN gkq_comm_s*& gkq_internal_comm_slot(gkq_handle_t h);
GKQ_HIDDEN int gkq_internal_device(gkq_handle_t h);
GKQ_HIDDEN unsigned long long gkq_internal_max_evals_seen(gkq_handle_t h);
GKQ_HIDDEN void gkq_internal_count_launch(gkq_handle_t h);
