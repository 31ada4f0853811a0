// sr_numa.h -- host-side placement helpers (sr_numa.cpp); private to the library.
// The end-to-end call is bound by the H2D copy of the caller's PCM, so where the pinned pages live and where the
// sender / packer threads run matters: a GPU hanging off socket 1 fed from memory on socket 0 shares the socket
// interconnect with every other rank doing the same (round 1: 36 ms instead of 19 ms per step at 8 GPUs).
#pragma once
#include <sched.h>
#include <stddef.h>
#include <stdint.h>

namespace srk {

// NUMA node of a PCI device given its bus id ("0000:1b:00.0", any case); -1 = unknown / single node
int numa_node_of_pci(const char *bus_id);
// CPUs of a node (sysfs cpulist) intersected with the calling thread's current affinity; false if empty / unknown
bool cpus_of_node(int node, cpu_set_t *out);
// number of NUMA nodes the kernel exposes (>= 1)
int numa_node_count();
// node that backs the page at p (get_mempolicy MPOL_F_NODE|MPOL_F_ADDR); -1 if the kernel will not say
int numa_node_of_page(const void *p);

// RAII: run the enclosing scope on the CPUs of `node` (no-op for node < 0 or when the node has no usable CPU),
// restoring the previous mask afterwards. Page faults taken inside the scope are served node-locally (first touch).
struct ScopedNodeAffinity {
    cpu_set_t prev;
    bool active = false;
    explicit ScopedNodeAffinity(int node);
    ~ScopedNodeAffinity();
};

// anonymous page-aligned memory whose pages are faulted in on `node` (first touch under ScopedNodeAffinity, plus an
// mbind preference when the kernel allows it); release with node_free. NULL on failure.
void *node_alloc(size_t bytes, int node);
void node_free(void *p, size_t bytes);

}  // namespace srk
