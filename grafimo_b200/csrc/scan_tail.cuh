// scan_tail.cuh -- what the host-buffer entry points (scan_host.cu, seqscan.cu) share: the device buffers of one scan
// and the steps after the last chunk (K5, K6, copy of the hit table to the host).  Not part of the ABI.
#pragma once
#include "internal.cuh"

struct gb2_scan_out {  // caller's host buffers (see gb2_scan_host in grafimo_b200.h)
    uint64_t *h_row;
    uint8_t *h_strand;
    int32_t *h_iscore;
    double *h_score, *h_p, *h_q;
    uint64_t *h_n_hits, *h_stats;
};

struct gb2_scan_bufs {
    uint64_t *d_hist, *d_cnt;  // d_cnt: [0]=N rows [1]=bad rows [2]=hit count [3]=kept [4]=total; follows d_hist in memory
    size_t hist_and_cnt_bytes;
    gb2_hit *d_hits;
    double *d_qtab;
    uint32_t *d_rank;
    uint64_t *o_row;
    double *o_score, *o_p, *o_q;
    int32_t *o_iscore;
    uint8_t *o_strand;
};

size_t gb2_scan_tail_bytes(const gb2_motif *m, uint64_t hit_capacity);
char *gb2_scan_tail_carve(char *q, const gb2_motif *m, uint64_t hit_capacity, gb2_scan_bufs &b);
int gb2_scan_tail_finish(gb2_ctx *ctx, const gb2_motif *m, const gb2_scan_bufs &b, uint64_t windows, uint64_t row_limit,
                         double p_threshold, int q_filter, int want_q, uint64_t hit_capacity, const gb2_scan_out &o);
int gb2_scan_check_args(gb2_ctx *ctx, const char *who, int strands, int q_filter, int want_q, uint64_t hit_capacity,
                        const gb2_scan_out &o);
