// Replay of long chains of short dependent kernels as CUDA graphs.
//
// The refit solvers issue hundreds to thousands of tiny kernels whose sequence depends only on the problem
// sizes and whose arguments depend only on the buffer addresses.  Issued one by one they are bound by the
// host's launch rate (~12 us per launch measured on the B200 hosts), not by the GPU.  sdb_graph_run captures
// the sequence once per key (buffer addresses + sizes) and replays the instantiated graph afterwards.
// The legacy default stream cannot be captured and a stream that is already being captured just records the
// kernels: in both cases the sequence is issued directly.  SDB_GRAPHS=0 (or SDB_LU_GRAPH=0) disables replay.
#pragma once
#include <stdlib.h>

#include "sdb_common.cuh"

struct sdb_graph_key {
    const void *p[5];
    int64_t n[4];
    int tag, device;
};

struct sdb_graph_entry {
    sdb_graph_key key;
    cudaGraphExec_t exec;
    unsigned long long last_use;
    long long nodes;
};

#define SDB_GRAPH_SLOTS 8

static inline bool sdb_graphs_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *a = getenv("SDB_GRAPHS"), *b = getenv("SDB_LU_GRAPH");
        v = ((a && a[0] == '0') || (b && b[0] == '0')) ? 0 : 1;
    }
    return v == 1;
}

static inline bool sdb_graph_key_eq(const sdb_graph_key &a, const sdb_graph_key &b) {
    for (int i = 0; i < 5; ++i)
        if (a.p[i] != b.p[i]) return false;
    for (int i = 0; i < 4; ++i)
        if (a.n[i] != b.n[i]) return false;
    return a.tag == b.tag && a.device == b.device;
}

// `enqueue` issues the kernel sequence on `st` and returns an SDB_* code.
template <class F>
static int sdb_graph_run(cudaStream_t st, sdb_graph_key key, F enqueue) {
    static sdb_graph_entry cache[SDB_GRAPH_SLOTS];
    static unsigned long long tick = 0;
    static bool warmed = false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (st != nullptr) cudaStreamIsCapturing(st, &cs);
    if (!sdb_graphs_enabled() || st == nullptr || cs != cudaStreamCaptureStatusNone) return enqueue();
    SDB_CUDA(cudaGetDevice(&key.device));
    sdb_graph_entry *hit = nullptr;
    for (int i = 0; i < SDB_GRAPH_SLOTS; ++i)
        if (cache[i].exec && sdb_graph_key_eq(cache[i].key, key)) { hit = &cache[i]; break; }
    if (!hit) {
        if (!warmed) {   // first call in the process: load modules / set attributes outside any capture
            warmed = true;
            return enqueue();
        }
        sdb_graph_entry *victim = nullptr;
        for (int i = 0; i < SDB_GRAPH_SLOTS && !victim; ++i)
            if (!cache[i].exec) victim = &cache[i];
        if (!victim) {
            victim = &cache[0];
            for (int i = 1; i < SDB_GRAPH_SLOTS; ++i)
                if (cache[i].last_use < victim->last_use) victim = &cache[i];
        }
        const long long before = sdb_launch_count(0);
        cudaGraph_t graph = nullptr;
        SDB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) { sdb_set_error("graph capture failed: %s", cudaGetErrorString(ce)); return SDB_ERR_CUDA; }
        if (victim->exec) cudaGraphExecDestroy(victim->exec);
        victim->exec = nullptr;
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) { sdb_set_error("graph instantiation failed: %s", cudaGetErrorString(ci)); return SDB_ERR_CUDA; }
        victim->key = key;
        victim->exec = exec;
        victim->nodes = sdb_launch_count(0) - before;
        hit = victim;
    } else {
        for (long long k = 0; k < hit->nodes; ++k) sdb_count_launch();   // replayed kernels are launches of this library too
    }
    hit->last_use = ++tick;
    SDB_CUDA(cudaGraphLaunch(hit->exec, st));
    return SDB_OK;
}
