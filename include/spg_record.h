/*
 * spg_record.h — sizes and offsets of the packed blanket records described in spg_capi.h.
 * Header-only, plain C, no dependencies: shared by the product library, by the CPU oracle and by
 * callers that pack records themselves. Units are 8-byte words.
 */
#ifndef SPG_RECORD_H_
#define SPG_RECORD_H_

#include <stdint.h>
#include "spg_capi.h"

#ifdef __CUDACC__
#define SPGR_FN static __host__ __device__ inline
#else
#define SPGR_FN static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

SPGR_FN int64_t spgr_pad2(int64_t n_int32) { return (n_int32 + 1) / 2; }
SPGR_FN int64_t spgr_pose_words(int32_t dim) { return dim == 3 ? 3 : 7; }

/* words of one edge inside an input record */
SPGR_FN int64_t spgr_edge_words(int32_t dim, int32_t kind, int32_t nv, int32_t rows) {
    int64_t w = 2 + spgr_pad2(nv);
    if(kind == SPG_EDGE_POSE) {
        w += spgr_pose_words(dim) + (int64_t) dim * dim;
    } else if(kind == SPG_EDGE_GLC) {
        w += (int64_t) dim * nv + (int64_t) rows * dim * nv;
    } else { /* SPG_EDGE_MULTI: nmeas = rows/dim; int32 pairs[2*nmeas]; meas[nmeas][P]; info[rows*rows] */
        int64_t nmeas = rows / dim;
        w += spgr_pad2(2 * nmeas) + nmeas * spgr_pose_words(dim) + (int64_t) rows * rows;
    }
    return w;
}

/* words of the fixed part of an input record (header, ids, poses, edge offset table) */
SPGR_FN int64_t spgr_record_fixed_words(int32_t dim, int32_t n_vert, int32_t n_edges) {
    return SPG_REC_HEADER_WORDS + spgr_pad2(n_vert) + (int64_t) n_vert * spgr_pose_words(dim) +
           spgr_pad2(n_edges);
}
SPGR_FN int64_t spgr_ids_off(void) { return SPG_REC_HEADER_WORDS; }
SPGR_FN int64_t spgr_poses_off(int32_t n_vert) { return SPG_REC_HEADER_WORDS + spgr_pad2(n_vert); }
SPGR_FN int64_t spgr_edgetab_off(int32_t dim, int32_t n_vert) {
    return spgr_poses_off(n_vert) + (int64_t) n_vert * spgr_pose_words(dim);
}

/* number of edges the chosen provider emits for n_kept kept vertices (upper bound for GLC, where
 * rank-0 edges are dropped). Follows pseudo_chow_liu.cpp:41-86 and topology_provider_glc.cpp:113-183. */
SPGR_FN int32_t spgr_out_edge_count(int32_t algorithm, int32_t topology, double chord_ratio,
                                          int32_t n_kept) {
    int32_t n = n_kept;
    if(algorithm == SPG_ALG_GLC) {
        if(n <= 0) return 0;
        if(n == 1 || topology == SPG_TOPO_DENSE) return 1;
        return n; /* root unary + (n-1) tree edges */
    }
    if(n < 2) return 0;
    if(n == 2) return 1;
    int32_t m = (int32_t) ((1 + chord_ratio) * (n - 1));
    int32_t all = n * (n - 1) / 2;
    switch(topology) {
    case SPG_TOPO_TREE: return n - 1;
    case SPG_TOPO_SUBGRAPH: return m >= all ? all : m;
    case SPG_TOPO_DENSE: return all;
    case SPG_TOPO_CLIQUEY_DENSE: return 1;
    default: return n - 1; /* CliqueySubgraph: at most n-1 correlated groups (entries) */
    }
}

/* words of one output slot */
SPGR_FN int64_t spgr_out_slot_words(int32_t dim, int32_t algorithm, int32_t topology, int32_t n_kept) {
    int64_t P = spgr_pose_words(dim);
    if(algorithm == SPG_ALG_GLC) {
        int64_t nvcap = (topology == SPG_TOPO_DENSE || n_kept == 1) ? n_kept : 2;
        int64_t c = (int64_t) dim * nvcap;
        return 1 + spgr_pad2(nvcap) + c + c * c;
    }
    if(topology == SPG_TOPO_CLIQUEY_DENSE || topology == SPG_TOPO_CLIQUEY_SUBGRAPH)
        return 0; /* correlated topologies: variable-size entries, see spgr_out_record_words */
    return 1 + P + (int64_t) dim * dim;
}

/* Correlated (cliquey) NFR topologies emit a SEQUENCE of variable-size entries behind the header, one per correlated
 * skeleton tree of the pattern (pseudo_chow_liu.cpp:198-251), in pattern order:
 *     w0 : int32 nmeas | int32 rows                   rows = dim * nmeas
 *     int32 ab[2*nmeas] (padded)                      kept-list indices (a, b) of every measurement, pattern order
 *     double meas[nmeas][P]                           Z = Xa^-1 Xb (setMeasurementFromState)
 *     double info[rows*rows]                          column-major (X of the closed form)
 * nmeas == 1 is a plain pose edge, nmeas > 1 a MultiEdgeCorrelated. The spanning tree has n_kept - 1 measurements in
 * total, so one entry with all of them bounds the information words and (5 + P) words per measurement bound the rest. */
SPGR_FN int64_t spgr_out_entry_words(int32_t dim, int32_t nmeas) {
    return 1 + spgr_pad2(2 * (int64_t) nmeas) + (int64_t) nmeas * spgr_pose_words(dim) + ((int64_t) dim * nmeas) * ((int64_t) dim * nmeas);
}

SPGR_FN int64_t spgr_out_record_words(int32_t dim, int32_t algorithm, int32_t topology,
                                            double chord_ratio, int32_t n_kept) {
    if(algorithm == SPG_ALG_NFR && (topology == SPG_TOPO_CLIQUEY_DENSE || topology == SPG_TOPO_CLIQUEY_SUBGRAPH)) {
        int64_t nm = n_kept > 1 ? n_kept - 1 : 0;
        int64_t wc = SPG_OUT_HEADER_WORDS + nm * (5 + spgr_pose_words(dim)) + (nm * dim) * (nm * dim);
        return (wc + 1) & ~(int64_t) 1;
    }
    int64_t w = SPG_OUT_HEADER_WORDS +
                (int64_t) spgr_out_edge_count(algorithm, topology, chord_ratio, n_kept) *
                        spgr_out_slot_words(dim, algorithm, topology, n_kept);
    return (w + 1) & ~(int64_t) 1; /* keep records 16-byte aligned */
}

#ifdef __cplusplus
}
#endif
#endif /* SPG_RECORD_H_ */
