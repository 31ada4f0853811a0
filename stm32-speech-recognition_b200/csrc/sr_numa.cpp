// sr_numa.cpp -- NUMA placement of the host side of sr_recognise_batch: which node a GPU hangs off, node-local
// pinned memory, node-local worker threads. Plain sysfs + sched_setaffinity + first touch: works inside containers
// where libnuma is absent and the mempolicy syscalls may be filtered (they are used when allowed, never required).
#include "sr_numa.h"
#include <ctype.h>
#include <dirent.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace srk {

static bool read_line(const char *path, char *buf, size_t cap) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, (int)cap, f) != nullptr;
    fclose(f);
    return ok;
}

int numa_node_of_pci(const char *bus_id) {
    if (!bus_id || !*bus_id) return -1;
    char id[64];
    size_t n = 0;
    for (const char *p = bus_id; *p && n + 1 < sizeof id; ++p) id[n++] = (char)tolower((unsigned char)*p);
    id[n] = 0;
    // CUDA prints "0000:1B:00.0" (or an 8-digit domain); sysfs uses a 4-digit lower-case domain
    const char *colon = strchr(id, ':');
    char path[160], line[64];
    if (colon && colon - id == 8) snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", id + 4);
    else snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", id);
    if (!read_line(path, line, sizeof line)) return -1;
    const int node = atoi(line);
    return node < 0 ? -1 : node;
}

int numa_node_count() {
    DIR *d = opendir("/sys/devices/system/node");
    if (!d) return 1;
    int n = 0;
    while (dirent *e = readdir(d))
        if (strncmp(e->d_name, "node", 4) == 0 && isdigit((unsigned char)e->d_name[4])) ++n;
    closedir(d);
    return n < 1 ? 1 : n;
}

bool cpus_of_node(int node, cpu_set_t *out) {
    if (node < 0 || !out) return false;
    char path[96], line[4096];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    if (!read_line(path, line, sizeof line)) return false;
    cpu_set_t cur, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof cur, &cur) != 0) return false;
    for (char *p = line; *p;) {                                    // "0-47,96-143"
        while (*p && !isdigit((unsigned char)*p)) ++p;
        if (!*p) break;
        long a = strtol(p, &p, 10), b = a;
        if (*p == '-') b = strtol(p + 1, &p, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET((int)c, &cur)) CPU_SET((int)c, &want);
    }
    if (CPU_COUNT(&want) == 0) return false;
    *out = want;
    return true;
}

int numa_node_of_page(const void *p) {
#ifdef SYS_get_mempolicy
    int node = -1;
    // MPOL_F_NODE (1) | MPOL_F_ADDR (2): node of the page that backs address p
    if (syscall(SYS_get_mempolicy, &node, nullptr, 0ul, p, 3ul) == 0) return node;
#endif
    return -1;
}

ScopedNodeAffinity::ScopedNodeAffinity(int node) {
    cpu_set_t want;
    if (node < 0 || numa_node_count() < 2 || !cpus_of_node(node, &want)) return;
    if (sched_getaffinity(0, sizeof prev, &prev) != 0) return;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return;
    active = true;
}
ScopedNodeAffinity::~ScopedNodeAffinity() {
    if (active) sched_setaffinity(0, sizeof prev, &prev);
}

void *node_alloc(size_t bytes, int node) {
    if (bytes == 0) return nullptr;
    const size_t page = 4096, len = (bytes + page - 1) / page * page;
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    madvise(p, len, MADV_HUGEPAGE);                                  // fewer TLB / IOMMU entries; advisory
#ifdef SYS_mbind
    if (node >= 0 && node < 1024 && numa_node_count() > 1) {       // MPOL_PREFERRED (1): a hint, never an allocation failure
        unsigned long mask[16] = {0};
        mask[node / 64] = 1ul << (node % 64);
        syscall(SYS_mbind, p, len, 1, mask, 1025ul, 0u);
    }
#endif
    {
        ScopedNodeAffinity bind(node);                              // first touch on the node's own CPUs
        volatile unsigned char *q = static_cast<volatile unsigned char *>(p);
        for (size_t i = 0; i < len; i += page) q[i] = 0;
    }
    return p;
}
void node_free(void *p, size_t bytes) {
    if (!p) return;
    const size_t page = 4096;
    munmap(p, (bytes + page - 1) / page * page);
}

}  // namespace srk
