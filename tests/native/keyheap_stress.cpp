/*
 * Test program: the key heap's lookup-or-insert (pgs_keyheap_intern,
 * pg_strom_b200/csrc/kern_textlib.cuh) under real concurrency on the CPU -
 * the device atomics are mapped to the compiler's __atomic builtins, the L2
 * loads to acquire loads - to check the claim / publish / wait protocol:
 * N threads intern strings drawn from one pool in different orders; every
 * thread must get the same word for the same string, different strings must
 * get different words, and the heap must hold every string exactly once.
 * Test infrastructure only.
 *
 *   keyheap_stress <threads> <pool> <rounds> <nslots> <heap_bytes>
 */
#include <algorithm>
#include <sched.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "pgstrom_kds.h"
#define DEVFN static inline
#define GPUPREAGG_INCOL_SLOT(colidx) 0
typedef struct { cl_bool value; bool isnull; } pg_bool_t;
typedef struct { cl_int value; bool isnull; } pg_int4_t;
#define devfunc_int_comp(x,y)   ((x) < (y) ? -1 : ((x) > (y) ? 1 : 0))
static inline void STROM_SET_ERROR(cl_int *p_error, cl_int errcode)
{ if (errcode > *p_error) *p_error = errcode; }
template <typename T> static inline T atomicAdd(T *p, T v)
{ return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
template <typename T> static inline T atomicCAS(T *p, T c, T v)
{ __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); return c; }
template <typename T> static inline T atomicExch(T *p, T v)
{ return __atomic_exchange_n(p, v, __ATOMIC_ACQ_REL); }
#define PGS_KEYHEAP_HOST_ATOMICS 1
#include "kern_textlib.cuh"

int main(int argc, char **argv)
{
    int nthreads = argc > 1 ? atoi(argv[1]) : 8;
    int npool = argc > 2 ? atoi(argv[2]) : 4000;
    int rounds = argc > 3 ? atoi(argv[3]) : 4;
    unsigned nslots = argc > 4 ? (unsigned)atoi(argv[4]) : (1u << 14);
    size_t heap_bytes = argc > 5 ? (size_t)atoll(argv[5]) : (size_t)8 << 20;
    std::vector<std::string> pool;
    uint64_t seed = 12345;
    auto rnd = [&seed]() { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(seed >> 33); };
    for (int i = 0; i < npool; i++)
    {
        int len = 8 + (int)(rnd() % 40);
        std::string s;
        char tag[24];
        snprintf(tag, sizeof(tag), "%08d", i);      /* distinct by construction */
        s = tag;
        while ((int)s.size() < len)
            s.push_back((char)(rnd() % 255 + 1));
        pool.push_back(s);
    }
    std::vector<cl_ulong> slots(2 * (size_t)nslots, 0);
    std::vector<unsigned char> heap(heap_bytes);
    cl_ulong used = 0;
    pgs_keyheap.slots = slots.data();
    pgs_keyheap.nslots = nslots;
    pgs_keyheap.heap = heap.data();
    pgs_keyheap.heap_bytes = heap_bytes;
    pgs_keyheap.heap_used = &used;
    pgs_keyheap.max_probe = 512;
    std::vector<std::vector<cl_ulong>> words(nthreads, std::vector<cl_ulong>(npool, 0));
    std::vector<int> bad(nthreads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            uint64_t s = 777 + 31 * t;
            for (int r = 0; r < rounds; r++)
                for (int k = 0; k < npool; k++)
                {
                    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                    /* all threads sweep the pool in roughly the same order so that
                     * they meet on fresh strings, with a little jitter */
                    int i = (k + (int)((s >> 40) % 7)) % npool;
                    bool ok = false;
                    cl_ulong w = pgs_keyheap_intern((const unsigned char *)pool[i].data(),
                                                    (cl_int)pool[i].size(), &ok);
                    if (!ok) { bad[t]++; continue; }
                    if (words[t][i] == 0) words[t][i] = w;
                    else if (words[t][i] != w) bad[t] += 1000000;
                }
        });
    for (auto &x : th) x.join();
    long nbad = 0, mismatch = 0, missing = 0;
    for (int t = 0; t < nthreads; t++) nbad += bad[t];
    std::vector<cl_ulong> all;
    size_t expect_used = 0;
    for (int i = 0; i < npool; i++)
    {
        cl_ulong w = 0;
        for (int t = 0; t < nthreads; t++)
        {
            if (words[t][i] == 0) continue;
            if (w == 0) w = words[t][i];
            else if (w != words[t][i]) mismatch++;
        }
        if (w == 0) { missing++; continue; }
        all.push_back(w);
        expect_used += 8 + ((pool[i].size() + 7) & ~(size_t)7);
        /* the heap entry is the string */
        size_t off = (size_t)(w & 0x00FFFFFFFFFFFFFFULL);
        uint64_t len;
        memcpy(&len, heap.data() + off, 8);
        if ((w >> 56) != 0x80 || len != pool[i].size() ||
            memcmp(heap.data() + off + 8, pool[i].data(), len) != 0)
            mismatch++;
    }
    std::vector<cl_ulong> sorted = all;
    std::sort(sorted.begin(), sorted.end());
    long dups = 0;
    for (size_t i = 1; i < sorted.size(); i++)
        if (sorted[i] == sorted[i - 1]) dups++;
    printf("{\"threads\": %d, \"pool\": %d, \"failed_lookups\": %ld, \"mismatch\": %ld, "
           "\"missing\": %ld, \"duplicate_words\": %ld, \"heap_used\": %llu, \"expect_used\": %zu}\n",
           nthreads, npool, nbad, mismatch, missing, dups, (unsigned long long)used, expect_used);
    return (nbad == 0 && mismatch == 0 && missing == 0 && dups == 0 && used == expect_used) ? 0 : 1;
}
