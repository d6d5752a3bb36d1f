/*
 * cuda_layer.cpp - the thin C-ABI CUDA layer of the GpuPreAgg path.
 *
 * Replaces, for this path, the reference's OpenCL dispatch stack:
 *   opencl_entry.c   (dlopen of the vendor runtime)      -> CUDA runtime API
 *   opencl_devinfo.c (device discovery, WG sizing)       -> pgs_cuda_init, occupancy API
 *   opencl_devprog.c (source keyed program cache, async
 *                     clBuildProgram)                    -> NVRTC for sm_100a + CRC32 cache
 *   opencl_serv.c    (bgworker, N pthreads, pinned shmem)-> streams inside the caller
 *   mqueue.c         (shared-memory message queues)      -> tickets + CUDA events
 *   gpupreagg.c:3009-4240 (clserv_process_gpupreagg: 4 buffer allocs, k+3 H2D
 *                     copies, 3..3+log^2 kernel launches, 2 D2H per chunk)
 *                                                        -> 1 H2D + 1 kernel per chunk,
 *                                                           state persistent in HBM,
 *                                                           1 flush kernel + D2H of the
 *                                                           groups at end of scan
 * There is no CPU fallback in here: without a CUDA device every device call
 * fails with StromError_ServerNotReady.
 */
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pgstrom_cuda.h"
#include "kern_shared.h"
#include "pgs_plan.h"

namespace pgs { extern thread_local std::string last_error; }
using pgs::last_error;

/* headers of the device runtime, embedded at build time (the reference's
 * Makefile:31-77 turns each opencl_*.h into a C string the same way) */
#include "kernel_headers.inc"

static void
set_error(const char *fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error = buf;
}

#define CUDA_CHECK(call)                                                \
    do {                                                                \
        cudaError_t __rc = (call);                                      \
        if (__rc != cudaSuccess)                                        \
        {                                                               \
            set_error("%s failed: %s (%s:%d)", #call,                   \
                      cudaGetErrorString(__rc), __FILE__, __LINE__);    \
            return StromError_CudaInternal;                             \
        }                                                               \
    } while (0)

/* ------------------------------------------------------------------
 * SHA-256 (FIPS 180-4): key of the program cache.  The reference keys its
 * cache by a CRC32 of the source and compares the full text on a hit
 * (opencl_devprog.c:580-659); the on-disk cache here cannot compare texts, so
 * the key is a cryptographic digest of everything the binary depends on and
 * the file carries the digest for verification.
 * ------------------------------------------------------------------ */
struct Sha256
{
    uint32_t    h[8];
    uint64_t    nbytes = 0;
    unsigned char buf[64];
    size_t      fill = 0;

    Sha256()
    {
        static const uint32_t init[8] = {
            0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
            0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19 };
        memcpy(h, init, sizeof(h));
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const unsigned char *p)
    {
        static const uint32_t K[64] = {
            0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
            0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
            0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
            0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
            0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
            0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
            0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
            0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2 };
        uint32_t w[64], v[8];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) |
                   ((uint32_t)p[4 * i + 2] << 8) | (uint32_t)p[4 * i + 3];
        for (int i = 16; i < 64; i++)
        {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        memcpy(v, h, sizeof(v));
        for (int i = 0; i < 64; i++)
        {
            uint32_t S1 = rotr(v[4], 6) ^ rotr(v[4], 11) ^ rotr(v[4], 25);
            uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
            uint32_t t1 = v[7] + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(v[0], 2) ^ rotr(v[0], 13) ^ rotr(v[0], 22);
            uint32_t mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
            uint32_t t2 = S0 + mj;
            v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = v[3] + t1;
            v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t1 + t2;
        }
        for (int i = 0; i < 8; i++)
            h[i] += v[i];
    }
    void update(const void *data, size_t len)
    {
        const unsigned char *p = (const unsigned char *)data;
        nbytes += len;
        while (len > 0)
        {
            size_t n = std::min(len, sizeof(buf) - fill);
            memcpy(buf + fill, p, n);
            fill += n; p += n; len -= n;
            if (fill == sizeof(buf))
            {
                block(buf);
                fill = 0;
            }
        }
    }
    /* fields are length-prefixed so that (a, bc) and (ab, c) differ */
    void field(const void *data, size_t len)
    {
        uint64_t n = len;
        update(&n, sizeof(n));
        update(data, len);
    }
    std::string hex()
    {
        uint64_t bits = nbytes * 8;
        unsigned char pad = 0x80;
        update(&pad, 1);
        pad = 0;
        while (fill != 56)
            update(&pad, 1);
        unsigned char lenb[8];
        for (int i = 0; i < 8; i++)
            lenb[i] = (unsigned char)(bits >> (56 - 8 * i));
        update(lenb, 8);
        char out[65];
        for (int i = 0; i < 8; i++)
            snprintf(out + 8 * i, 9, "%08x", h[i]);
        return std::string(out, 64);
    }
};

/* ------------------------------------------------------------------
 * device programs
 * ------------------------------------------------------------------ */
struct pgs_program
{
    std::string key;        /* SHA-256 (hex) of source, flags, options, runtime headers, NVRTC */
    std::string source;
    int         extra_flags;
    std::string options;
    std::vector<char> cubin;
    std::string build_log;
    int         refcnt;
    double      build_ms;
    bool        from_disk;
};

static std::mutex program_lock;
static std::map<std::string, pgs_program *> program_cache;

/*
 * Where built programs are kept between processes: PGSTROM_CUBIN_CACHE, else
 * _cubin_cache/ next to the library (in-tree, so that pre-built programs
 * travel with it), else ~/.cache/pgstrom_cubin.  Never a shared directory
 * like /tmp: a cubin found there runs on the GPU with the caller's data.  The
 * directory is created 0700 and used only if it belongs to this user and
 * nobody else may write to it; otherwise there is no disk cache ("").
 */
static bool
cache_dir_usable(const std::string &dir)
{
    struct stat st;

    if (mkdir(dir.c_str(), 0700) != 0 && errno != EEXIST)
        return false;
    if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
        return false;
    if (st.st_uid != geteuid() || (st.st_mode & (S_IWGRP | S_IWOTH)) != 0)
        return false;
    return true;
}

static std::string
cubin_cache_dir()
{
    const char *env = getenv("PGSTROM_CUBIN_CACHE");
    std::vector<std::string> cands;
    if (env && *env)
        cands.push_back(env);
    else
    {
        Dl_info info;
        if (dladdr((void *)&cache_dir_usable, &info) && info.dli_fname)
        {
            std::string p = info.dli_fname;
            size_t slash = p.rfind('/');
            if (slash != std::string::npos)
                cands.push_back(p.substr(0, slash) + "/_cubin_cache");
        }
        const char *home = getenv("HOME");
        if (home && *home)
        {
            std::string c = std::string(home) + "/.cache";
            mkdir(c.c_str(), 0700);
            cands.push_back(c + "/pgstrom_cubin");
        }
    }
    for (auto &d : cands)
        if (cache_dir_usable(d))
            return d;
    return "";
}

/* file = magic, 64 hex digits of the key, cubin.  A file whose key is not
 * the one asked for (renamed, truncated, planted) is ignored. */
static const char CUBIN_FILE_MAGIC[8] = { 'P', 'G', 'S', 'C', 'U', 'B', '0', '2' };

static bool
cubin_file_read(const std::string &path, const std::string &key, std::vector<char> &cubin)
{
    FILE *fp = fopen(path.c_str(), "rb");
    char head[8 + 64];
    bool ok = false;

    if (!fp)
        return false;
    if (fread(head, 1, sizeof(head), fp) == sizeof(head) &&
        memcmp(head, CUBIN_FILE_MAGIC, 8) == 0 && memcmp(head + 8, key.data(), 64) == 0)
    {
        long pos = ftell(fp);
        fseek(fp, 0, SEEK_END);
        long sz = ftell(fp) - pos;
        fseek(fp, pos, SEEK_SET);
        if (sz > 0)
        {
            cubin.resize((size_t)sz);
            ok = (fread(cubin.data(), 1, (size_t)sz, fp) == (size_t)sz);
            if (!ok)
                cubin.clear();
        }
    }
    fclose(fp);
    return ok;
}

static void
cubin_file_write(const std::string &path, const std::string &key, const std::vector<char> &cubin)
{
    std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
    if (fd < 0)
        return;
    FILE *fp = fdopen(fd, "wb");
    if (!fp)
    {
        close(fd);
        unlink(tmp.c_str());
        return;
    }
    bool ok = fwrite(CUBIN_FILE_MAGIC, 1, 8, fp) == 8 &&
              fwrite(key.data(), 1, 64, fp) == 64 &&
              fwrite(cubin.data(), 1, cubin.size(), fp) == cubin.size();
    ok = (fclose(fp) == 0) && ok;
    if (ok)
        rename(tmp.c_str(), path.c_str());
    else
        unlink(tmp.c_str());
}

/*
 * NVRTC is resolved with dlopen, by absolute path first (the reference binds
 * its vendor runtime the same way, opencl_entry.c).  A process that has
 * another libnvrtc.so.12 mapped already - PyTorch preloads its bundled 12.8 -
 * would otherwise hand us that copy through the SONAME, and the device code
 * (256-bit global stores) needs the toolkit's 12.9 or later.  A path that
 * does not match a loaded SONAME string is opened as its own object.
 */
static struct nvrtc_entry
{
    void       *handle;
    std::string path;
    nvrtcResult (*Version)(int *, int *);
    const char *(*GetErrorString)(nvrtcResult);
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int,
                                 const char *const *, const char *const *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
} nvrtc;
static std::once_flag nvrtc_once;

static bool
nvrtc_bind(void *h)
{
#define NVRTC_SYM(name) \
    if (!(*(void **)(&nvrtc.name) = dlsym(h, "nvrtc" #name))) return false
    NVRTC_SYM(Version);
    NVRTC_SYM(GetErrorString);
    NVRTC_SYM(CreateProgram);
    NVRTC_SYM(CompileProgram);
    NVRTC_SYM(GetProgramLogSize);
    NVRTC_SYM(GetProgramLog);
    NVRTC_SYM(GetCUBINSize);
    NVRTC_SYM(GetCUBIN);
    NVRTC_SYM(DestroyProgram);
#undef NVRTC_SYM
    return true;
}

static void
nvrtc_load_once(void)
{
    std::vector<std::string> cands;
    const char *e;
    if ((e = getenv("PGSTROM_NVRTC_PATH")) && *e)
        cands.push_back(e);
    for (const char *var : { "CUDA_HOME", "CUDA_PATH" })
        if ((e = getenv(var)) && *e)
        {
            cands.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
            cands.push_back(std::string(e) + "/lib64/libnvrtc.so");
        }
    cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    cands.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    cands.push_back("libnvrtc.so.12");
    cands.push_back("libnvrtc.so");
    void       *fallback = NULL;
    std::string fallback_path;
    for (auto &c : cands)
    {
        void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        int   major = 0, minor = 0;
        if (!h)
            continue;
        if (!nvrtc_bind(h) || nvrtc.Version(&major, &minor) != NVRTC_SUCCESS)
        {
            dlclose(h);
            continue;
        }
        if (major > 12 || (major == 12 && minor >= 9))
        {
            nvrtc.handle = h;
            nvrtc.path = c;
            return;
        }
        if (!fallback)
        {
            fallback = h;
            fallback_path = c;
        }
    }
    /* an older compiler: the device code falls back to 128-bit stores */
    if (fallback && nvrtc_bind(fallback))
    {
        nvrtc.handle = fallback;
        nvrtc.path = fallback_path;
    }
}

static bool
nvrtc_ready(void)
{
    std::call_once(nvrtc_once, nvrtc_load_once);
    if (!nvrtc.handle)
        set_error("NVRTC not found (tried $PGSTROM_NVRTC_PATH, $CUDA_HOME/lib64, "
                  "/usr/local/cuda/lib64, the loader path): %s", dlerror());
    return nvrtc.handle != NULL;
}

static std::string
program_options(int extra_flags)
{
    std::string o;
    (void)extra_flags;
    o += "warps=" + std::string(getenv("PGSTROM_CONSUMER_WARPS") ? getenv("PGSTROM_CONSUMER_WARPS") : "16");
    o += ";minctas=" + std::string(getenv("PGSTROM_MIN_CTAS") ? getenv("PGSTROM_MIN_CTAS") : "1");
    if (getenv("PGSTROM_DEAL_STEPS"))
        o += ";dealsteps=" + std::string(getenv("PGSTROM_DEAL_STEPS"));
    o += ";opt=" + std::string(pgs::guc_bool("pg_strom.devprog_enable_optimization") ? "1" : "0");
    if (getenv("PGSTROM_DEBUG_LEVEL"))      /* timing experiments only: wrong results */
        o += ";debug=" + std::string(getenv("PGSTROM_DEBUG_LEVEL"));
    return o;
}

static int
nvrtc_build(pgs_program *prog)
{
    nvrtcProgram nprog;
    const char *headers[] = { pgs_hdr_pgstrom_kds_h, pgs_hdr_kern_shared_h,
                              pgs_hdr_kern_common_cuh, pgs_hdr_kern_mathlib_cuh,
                              pgs_hdr_kern_numeric_cuh, pgs_hdr_kern_timelib_cuh, pgs_hdr_kern_textlib_cuh,
                              pgs_hdr_kern_gpupreagg_cuh };
    const char *names[] = { "pgstrom_kds.h", "kern_shared.h", "kern_common.cuh",
                            "kern_mathlib.cuh", "kern_numeric.cuh", "kern_timelib.cuh", "kern_textlib.cuh",
                            "kern_gpupreagg.cuh" };
    std::string d_warps = "-DGPUPREAGG_CONSUMER_WARPS=" +
        std::string(getenv("PGSTROM_CONSUMER_WARPS") ? getenv("PGSTROM_CONSUMER_WARPS") : "16");
    std::string d_rpt = "-DGPUPREAGG_MIN_CTAS=" +
        std::string(getenv("PGSTROM_MIN_CTAS") ? getenv("PGSTROM_MIN_CTAS") : "1");
    std::string d_dbg = "-DGPUPREAGG_DEBUG_LEVEL=" +
        std::string(getenv("PGSTROM_DEBUG_LEVEL") ? getenv("PGSTROM_DEBUG_LEVEL") : "0");
    std::string d_deal = "-DGPUPREAGG_DEAL_STEPS=" +
        std::to_string(std::min(4, std::max(1, atoi(getenv("PGSTROM_DEAL_STEPS")
                                                     ? getenv("PGSTROM_DEAL_STEPS") : "1"))));
    std::vector<const char *> opts = {
        "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
        "-device-int128", "--fmad=false",
        d_warps.c_str(), d_rpt.c_str(), d_dbg.c_str(),
    };
    if (getenv("PGSTROM_DEAL_STEPS"))
        opts.push_back(d_deal.c_str());
    if (!pgs::guc_bool("pg_strom.devprog_enable_optimization"))
        opts.push_back("-Xptxas=-O0");
    if (!nvrtc_ready())
        return StromError_ProgramBuildFailure;
    auto t0 = std::chrono::steady_clock::now();
    nvrtcResult rc = nvrtc.CreateProgram(&nprog, prog->source.c_str(), "gpupreagg.cu",
                                        (int)(sizeof(headers) / sizeof(headers[0])),
                                        headers, names);
    if (rc != NVRTC_SUCCESS)
    {
        set_error("nvrtcCreateProgram: %s", nvrtc.GetErrorString(rc));
        return StromError_ProgramBuildFailure;
    }
    rc = nvrtc.CompileProgram(nprog, (int)opts.size(), opts.data());
    size_t logsz = 0;
    nvrtc.GetProgramLogSize(nprog, &logsz);
    prog->build_log.assign(logsz, '\0');
    if (logsz > 1)
        nvrtc.GetProgramLog(nprog, &prog->build_log[0]);
    if (rc != NVRTC_SUCCESS)
    {
        set_error("device program build failure: %s", nvrtc.GetErrorString(rc));
        nvrtc.DestroyProgram(&nprog);
        return StromError_ProgramBuildFailure;
    }
    size_t sz = 0;
    nvrtc.GetCUBINSize(nprog, &sz);
    prog->cubin.resize(sz);
    nvrtc.GetCUBIN(nprog, prog->cubin.data());
    nvrtc.DestroyProgram(&nprog);
    prog->build_ms = std::chrono::duration<double, std::milli>(
        std::chrono::steady_clock::now() - t0).count();
    return StromError_Success;
}

extern "C" int
pgs_program_build(const char *kern_source, int extra_flags,
                  pgs_program **program, const char **build_log)
{
    static thread_local std::string log_buf;
    std::string options = program_options(extra_flags);
    std::string key;

    if (build_log)
        *build_log = NULL;
    if (!kern_source || !program)
    {
        set_error("pgs_program_build: bad arguments");
        return StromError_BadRequestMessage;
    }
    if (!nvrtc_ready())
        return StromError_ProgramBuildFailure;
    /* everything the binary depends on: the query's source and flags, the
     * build options, the static device runtime (a new library build must not
     * pick up stale binaries) and the compiler itself */
    {
        Sha256 sha;
        int nv_major = 0, nv_minor = 0;
        nvrtc.Version(&nv_major, &nv_minor);
        sha.field(kern_source, strlen(kern_source));
        sha.field(&extra_flags, sizeof(extra_flags));
        sha.field(options.data(), options.size());
        sha.field(&nv_major, sizeof(nv_major));
        sha.field(&nv_minor, sizeof(nv_minor));
        sha.field(pgs_hdr_kern_gpupreagg_cuh, strlen(pgs_hdr_kern_gpupreagg_cuh));
        sha.field(pgs_hdr_kern_common_cuh, strlen(pgs_hdr_kern_common_cuh));
        sha.field(pgs_hdr_kern_mathlib_cuh, strlen(pgs_hdr_kern_mathlib_cuh));
        sha.field(pgs_hdr_kern_numeric_cuh, strlen(pgs_hdr_kern_numeric_cuh));
        sha.field(pgs_hdr_kern_timelib_cuh, strlen(pgs_hdr_kern_timelib_cuh));
        sha.field(pgs_hdr_kern_textlib_cuh, strlen(pgs_hdr_kern_textlib_cuh));
        sha.field(pgs_hdr_pgstrom_kds_h, strlen(pgs_hdr_pgstrom_kds_h));
        sha.field(pgs_hdr_kern_shared_h, strlen(pgs_hdr_kern_shared_h));
        key = sha.hex();
    }

    std::lock_guard<std::mutex> g(program_lock);
    auto it = program_cache.find(key);
    if (it != program_cache.end())
    {
        it->second->refcnt++;
        *program = it->second;
        return StromError_Success;
    }
    pgs_program *prog = new pgs_program;
    prog->key = key;
    prog->source = kern_source;
    prog->extra_flags = extra_flags;
    prog->options = options;
    prog->refcnt = 1;
    prog->build_ms = 0;
    prog->from_disk = false;

    /* on-disk cache (in-tree by default so that pre-built programs travel) */
    std::string dir = cubin_cache_dir();
    std::string path = dir.empty() ? "" : dir + "/" + key.substr(0, 40) + ".cubin";
    if (!path.empty() && cubin_file_read(path, key, prog->cubin))
        prog->from_disk = true;
    if (prog->cubin.empty())
    {
        int rc = nvrtc_build(prog);
        if (rc != StromError_Success)
        {
            log_buf = prog->build_log + "\n---- source ----\n" + prog->source;
            if (build_log)
                *build_log = log_buf.c_str();
            delete prog;
            return rc;
        }
        if (!path.empty())
            cubin_file_write(path, key, prog->cubin);
    }
    if (build_log && !prog->build_log.empty())
    {
        log_buf = prog->build_log;
        *build_log = log_buf.c_str();
    }
    program_cache[key] = prog;
    *program = prog;
    return StromError_Success;
}

extern "C" void
pgs_program_release(pgs_program *program)
{
    std::lock_guard<std::mutex> g(program_lock);
    if (program && program->refcnt > 0)
        program->refcnt--;
    /* kept in the cache; reclaimed by size (devprog_reclaim_threshold) */
    size_t total = 0;
    for (auto &kv : program_cache)
        total += kv.second->cubin.size() + kv.second->source.size();
    size_t limit = (size_t)pgs::guc_int("pg_strom.devprog_reclaim_threshold") << 10;
    if (total > limit)
    {
        for (auto it = program_cache.begin(); it != program_cache.end() && total > limit; )
        {
            if (it->second->refcnt == 0)
            {
                total -= it->second->cubin.size() + it->second->source.size();
                delete it->second;
                it = program_cache.erase(it);
            }
            else
                ++it;
        }
    }
}

extern "C" const void *
pgs_program_cubin(pgs_program *program, size_t *length)
{
    if (!program)
        return NULL;
    if (length)
        *length = program->cubin.size();
    return program->cubin.data();
}

extern "C" const char *
pgs_program_info_json(void)
{
    static thread_local std::string buf;
    std::lock_guard<std::mutex> g(program_lock);
    pgs::JsonPtr arr = pgs::Json::array();
    for (auto &kv : program_cache)
    {
        pgs::JsonPtr o = pgs::Json::object();
        o->set("key", kv.first.substr(0, 16));
        o->set("refcnt", kv.second->refcnt);
        o->set("length", (long long)kv.second->source.size());
        o->set("cubin_length", (long long)kv.second->cubin.size());
        o->set("build_ms", pgs::Json::number(kv.second->build_ms));
        o->setb("from_disk", kv.second->from_disk);
        arr->push(o);
    }
    buf = arr->dump();
    return buf.c_str();
}

/* ------------------------------------------------------------------
 * devices
 * ------------------------------------------------------------------ */
struct DeviceInfo
{
    int             ordinal;
    cudaDeviceProp  prop;
};
static std::vector<DeviceInfo> devices;
static std::mutex device_lock;

extern "C" int
pgs_cuda_init(const int *devs, int ndevices)
{
    std::lock_guard<std::mutex> g(device_lock);
    int count = 0;
    cudaError_t rc = cudaGetDeviceCount(&count);

    if (rc != cudaSuccess || count == 0)
    {
        set_error("no CUDA device available: %s",
                  rc != cudaSuccess ? cudaGetErrorString(rc) : "device count is 0");
        devices.clear();
        return StromError_ServerNotReady;
    }
    devices.clear();
    std::vector<int> list;
    if (devs && ndevices > 0)
        list.assign(devs, devs + ndevices);
    else
    {
        std::string guc = pgs::guc_get("pg_strom.opencl_devices");
        if (guc.empty() || guc == "any")
            for (int i = 0; i < count; i++) list.push_back(i);
        else
        {
            size_t pos = 0;
            while (pos < guc.size())
            {
                size_t comma = guc.find(',', pos);
                if (comma == std::string::npos) comma = guc.size();
                list.push_back(atoi(guc.substr(pos, comma - pos).c_str()));
                pos = comma + 1;
            }
        }
    }
    for (int ord : list)
    {
        DeviceInfo di;
        if (ord < 0 || ord >= count)
        {
            set_error("CUDA device %d does not exist (%d devices)", ord, count);
            devices.clear();
            return StromError_BadRequestMessage;
        }
        di.ordinal = ord;
        CUDA_CHECK(cudaGetDeviceProperties(&di.prop, ord));
        if (di.prop.major < 10)
        {
            set_error("CUDA device %d (%s, sm_%d%d) is not a Blackwell sm_100 device; "
                      "this library carries sm_100a code only",
                      ord, di.prop.name, di.prop.major, di.prop.minor);
            devices.clear();
            return StromError_ServerNotReady;
        }
        devices.push_back(di);
    }
    return StromError_Success;
}

extern "C" int
pgs_cuda_device_count(void)
{
    return (int)devices.size();
}

extern "C" const char *
pgs_cuda_device_info_json(void)
{
    static thread_local std::string buf;
    pgs::JsonPtr arr = pgs::Json::array();
    for (auto &d : devices)
    {
        pgs::JsonPtr o = pgs::Json::object();
        o->set("ordinal", d.ordinal);
        o->set("name", d.prop.name);
        o->set("sm_count", d.prop.multiProcessorCount);
        o->set("compute_capability", std::to_string(d.prop.major) + "." + std::to_string(d.prop.minor));
        o->set("global_mem_bytes", (long long)d.prop.totalGlobalMem);
        o->set("l2_bytes", (long long)d.prop.l2CacheSize);
        o->set("smem_per_block_optin", (long long)d.prop.sharedMemPerBlockOptin);
        arr->push(o);
    }
    buf = arr->dump();
    return buf.c_str();
}

extern "C" void
pgs_cuda_shutdown(void)
{
    std::lock_guard<std::mutex> g(device_lock);
    devices.clear();
}

static int
device_ordinal(int index)
{
    if (index < 0 || (size_t)index >= devices.size())
        return -1;
    return devices[index].ordinal;
}

extern "C" void *
pgs_chunk_alloc(size_t length)
{
    void *p = NULL;
    if (devices.empty())
    {
        /* no device: plain aligned memory so that host-only callers
         * (planner tests) can still build chunks */
        if (posix_memalign(&p, 4096, (length + 4095) & ~(size_t)4095) != 0)
            return NULL;
        return p;
    }
    if (cudaHostAlloc(&p, length, cudaHostAllocPortable) != cudaSuccess)
    {
        set_error("cudaHostAlloc(%zu) failed", length);
        return NULL;
    }
    return p;
}

extern "C" void
pgs_chunk_free(void *chunk)
{
    if (!chunk)
        return;
    if (devices.empty())
    {
        free(chunk);
        return;
    }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, chunk) == cudaSuccess &&
        attr.type == cudaMemoryTypeHost)
        cudaFreeHost(chunk);
    else
    {
        cudaGetLastError();
        free(chunk);
    }
}

extern "C" void *
pgs_device_alloc(int device, size_t length)
{
    void *p = NULL;
    int ord = device_ordinal(device);
    if (ord < 0 || cudaSetDevice(ord) != cudaSuccess ||
        cudaMalloc(&p, length) != cudaSuccess)
    {
        set_error("pgs_device_alloc(%zu) failed: %s", length,
                  cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    return p;
}

extern "C" void
pgs_device_free(int device, void *ptr)
{
    int ord = device_ordinal(device);
    if (ord >= 0 && cudaSetDevice(ord) == cudaSuccess)
        cudaFree(ptr);
}

extern "C" int
pgs_device_upload(int device, void *dst_device, const void *src_host, size_t length)
{
    int ord = device_ordinal(device);
    if (ord < 0)
    {
        set_error("device layer is not initialised");
        return StromError_ServerNotReady;
    }
    CUDA_CHECK(cudaSetDevice(ord));
    CUDA_CHECK(cudaMemcpy(dst_device, src_host, length, cudaMemcpyHostToDevice));
    return StromError_Success;
}

/* ------------------------------------------------------------------
 * sessions
 * ------------------------------------------------------------------ */
/* block size of the heap-page kernel: small blocks, so that the register file
 * (not the block shape of the staged kernel) decides how many warps an SM holds */
#define PGS_HEAP_BLOCK_THREADS  256

struct ChunkResult
{
    int32_t     status = 0;
    bool        done = false;
    std::vector<uint32_t> recheck_rows;
};

struct ChunkSlot
{
    void       *d_kds = NULL;       /* staging copy of the chunk in HBM */
    size_t      d_kds_cap = 0;
    char       *d_kgpreagg = NULL;  /* kern_gpupreagg (status, kparams, row map) */
    char       *h_kgpreagg = NULL;  /* pinned image */
    size_t      kg_cap = 0;
    cl_uint    *d_heap_index = NULL;/* heap pages: first row item of every page */
    cl_uint    *d_recheck = NULL;   /* 1 bit per row */
    size_t      recheck_words = 0;
    int32_t    *h_status = NULL;    /* pinned */
    cudaEvent_t ev_copied = NULL;
    cudaEvent_t ev_done = NULL;
    cudaEvent_t ev_k0 = NULL, ev_k1 = NULL; /* around the main kernel (perfmon) */
    bool        timed = false;
    uint64_t    table_epoch = 0;    /* num_table_grown when the chunk was launched */
    pgs_ticket  ticket = -1;
    uint32_t    nitems = 0;
    bool        busy = false;
};

struct pgs_session
{
    int             device_index = 0;
    int             ordinal = 0;
    pgs_program    *program = NULL;
    cudaLibrary_t   library = NULL;
    cudaKernel_t    k_main = NULL, k_rowmap = NULL, k_heap = NULL, k_partagg = NULL,
                    k_heap_index = NULL, k_heap_staged = NULL,
                    k_init = NULL, k_flush = NULL, k_rehash = NULL,
                    k_export = NULL, k_import = NULL, k_import_blocks = NULL,
                    k_export_parts = NULL, k_peer_push = NULL, k_peer_pull = NULL,
                    k_describe = NULL;
    pgs_kern_desc   desc;
    pgs_gstate      gs;
    std::vector<char> kparams;
    pgs_session_config config;
    std::vector<kern_colmeta> result_colmeta;
    cudaStream_t    s_copy = NULL, s_exec = NULL;
    std::vector<ChunkSlot> slots;
    pgs_ticket      next_ticket = 0;
    std::map<pgs_ticket, ChunkResult> results;
    int             grid_main = 0;
    size_t          smem_main = 0;
    cl_uint         sh_nslots = 0;
    int             grid_heap = 0;      /* heap-page kernel: no staging ring, more CTAs per SM */
    size_t          smem_heap = 0;
    cl_uint         heap_pps = 0;       /* staged heap kernel: pages per stage (0 = off) */
    cl_uint         heap_nstages = 0;
    size_t          smem_heap_staged = 0;
    int             grid_partagg = 0;
    size_t          smem_partagg = 0;
    cl_uint         tile_rows = 2048;
    cl_uint         nstages = 4;
    int             num_sms = 0;
    uint64_t        launches = 0;
    /* perfmon (pg_strom.h:177-213) */
    uint64_t        num_dma_send = 0, bytes_dma_send = 0;
    uint64_t        num_dma_recv = 0, bytes_dma_recv = 0;
    uint64_t        num_chunks = 0, num_rechecked_chunks = 0;
    double          time_kern_build_ms = 0;
    bool            perfmon = false;
    double          time_kern_main_ms = 0;  /* sum over launches, from CUDA events */
    uint64_t        num_kern_main = 0;
    uint64_t        rows_kern_main = 0;
    void           *d_scratch = NULL;   /* small device buffer: desc, counters */
    char           *d_result = NULL;    /* TUPSLOT store written by the flush kernel */
    size_t          d_result_cap = 0;
    kern_gpupreagg *d_kg_misc = NULL;   /* status word of flush / import kernels */
    char           *h_result_head = NULL;   /* pinned: header + status read-back */
    char           *d_xchg = NULL;          /* NCCL merge: send block + receive blocks */
    size_t          d_xchg_cap = 0;
    char           *d_xrecv = NULL;         /* partitioned exchange: records received */
    size_t          d_xrecv_cap = 0;
    int             ovf_which = 0;          /* which of the two overflow counters is live */
    std::vector<void *> graveyard;          /* tables replaced by larger ones: freed once
                                             * nothing in flight can still read them */
    uint64_t        num_table_grown = 0;
    /* merge of small states over NVLink peer memory (pgs_preagg_peer_*) */
    cl_ulong       *peer_area = NULL;       /* this rank's exchange area */
    size_t          peer_area_bytes = 0;
    cl_ulong       *peer_root_area = NULL;  /* the root's, mapped over NVLink (or peer_area) */
    bool            peer_mapped = false;
    int             peer_rank = -1, peer_nranks = 0, peer_root = 0;
    cl_uint         peer_cap = 0;
    cl_ulong        peer_epoch = 0;
    /* merge trace (perfmon): CUDA events around the phases of the last merges */
    cudaEvent_t     ev_m[4] = {NULL, NULL, NULL, NULL};
    double          merge_ms[3] = {0, 0, 0};
    uint64_t        merge_count = 0;
    bool            merge_pending = false;  /* d_kg_misc carries the status of a merge */
    size_t          seg_cursor_bytes = 0;   /* segment mode: shared memory of the cursors */
    /* a rank that pushed its whole state to the root has nothing to flush:
     * the push kernel says so in a word of mapped pinned memory */
    cl_uint        *h_moved = NULL;
    bool            push_pending = false;
    std::string     perfmon_buf;
    bool            aborted = false;
    /* key heap: long text / bpchar grouping keys (kern_textlib.cuh) */
    pgs_keyheap_ctl kh = {NULL, NULL, NULL, 0, 0, 0};
    std::vector<unsigned char> kh_host;     /* the heap as the last finish saw it */
    uint64_t        kh_finishes = 0;
};

static int launch_kernel(pgs_session *s, cudaKernel_t k, int grid, int block,
                         size_t smem, void **args);

/* Partitioned GROUP BY (very many groups): how many partitions, how many
 * slots per table image.  0 partitions = not partitioned. */
static size_t
partition_plan(const pgs_session *s, size_t *p_slots)
{
    if (!(s->desc.num_keys > 0 && s->desc.part_rec_bytes > 0 &&
          s->config.num_groups >= 65536.0 && !getenv("PGSTROM_NO_PARTITION")))
        return 0;
    const char *es = getenv("PGSTROM_PART_SLOTS");
    size_t slots = es ? (((size_t)atol(es) + 31) & ~(size_t)31) : 1024;
    size_t smem_max = devices[s->config.device].prop.sharedMemPerBlockOptin;
    size_t head = std::max<size_t>(128, s->desc.partagg_head_bytes);
    while (slots > 64 && head + slots * s->desc.sh_slot_bytes > smem_max)
        slots -= 32;
    if (p_slots)
        *p_slots = slots;
    return (size_t)(s->config.num_groups / (slots * 0.5)) + 1;
}

static int
session_alloc_state(pgs_session *s)
{
    size_t W = 1 + s->desc.num_cells;
    size_t max_ctas = (size_t)s->num_sms * 16;
    char   *base;
    size_t  sz_small = 4096;

    CUDA_CHECK(cudaMalloc(&s->d_scratch, sz_small + 8 * W * (1 + max_ctas)));
    CUDA_CHECK(cudaMemset(s->d_scratch, 0, sz_small + 8 * W * (1 + max_ctas)));
    base = (char *)s->d_scratch;
    /* words of the scratch page: 0 ticket | 8 ngroups | 16 rows scanned |
     * 24 rows filtered | 32 table occupancy | 40, 48 overflow counters |
     * 64 export counter | 80 peer push / pull counters | 128.. debug |
     * 512.. exchange counts */
    s->gs.ng_ticket = (cl_uint *)(base + 0);
    s->gs.gh_ngroups = (cl_uint *)(base + 8);
    s->gs.nrows_scanned = (cl_ulong *)(base + 16);
    s->gs.nrows_filtered = (cl_ulong *)(base + 24);
    s->gs.gh_nused = (cl_uint *)(base + 32);
    s->gs.ovf_count = (cl_uint *)(base + 40);
    s->gs.ovf_recs = NULL;
    s->gs.ovf_cap = 0;
    s->gs.ovf_pad = 0;
    s->ovf_which = 0;
    s->gs.ng_state = (cl_ulong *)(base + sz_small);
    s->gs.ng_partial = s->gs.ng_state + W;
    s->gs.gh_slots = NULL;
    s->gs.gh_nslots = 0;
    s->gs.gh_max_probe = 0;
    s->gs.part_nparts = 0;
    s->gs.part_cap = 0;
    s->gs.part_slots = 0;
    s->gs.part_pad = 0;
    s->gs.part_cursor = NULL;
    s->gs.part_nused = NULL;
    s->gs.part_recs = NULL;
    s->gs.part_images = NULL;
    s->grid_partagg = 0;
    s->smem_partagg = 0;
    s->gs.part_seg_cap = 0;
    s->gs.part_seg_max = 0;
    s->gs.part_seg_counts = NULL;
    size_t plan_slots = 0;
    size_t plan_nparts = (s->sh_nslots == 0 ? partition_plan(s, &plan_slots) : 0);
    if (plan_nparts != 0)
    {
        /* very many groups: partitioned aggregation (gpupreagg_partagg).
         * An image holds `slots` groups at most 75% full; partitions are
         * sized for half of that on average, their record areas for the
         * rows of one chunk plus a quarter (segment mode: plus a half,
         * segment by segment, + 16). */
        size_t slots = plan_slots;
        size_t image_bytes = slots * s->desc.sh_slot_bytes;
        size_t nparts = plan_nparts;
        size_t max_rows = s->config.max_chunk_rows ? s->config.max_chunk_rows : (64u << 20);
        size_t cap = ((size_t)((double)max_rows / nparts * 1.25) + 64 + 3) & ~(size_t)3;
        size_t seg_cap = 0, seg_max = 0;
        if (s->seg_cursor_bytes != 0)
        {
            /* one segment per CTA of the scan kernel, cursors in its shared
             * memory (see pgs_part_reserve_seg for the 16-bit bound) */
            seg_max = (size_t)s->grid_main;
            seg_cap = ((size_t)((double)max_rows / nparts / seg_max * 1.5) + 16 + 3) & ~(size_t)3;
            if (seg_max > 256 || seg_cap + 8 * (size_t)s->desc.block_threads >= 65535)
                seg_cap = seg_max = 0;
            else
                cap = seg_cap * seg_max;
        }
        size_t rec_bytes = nparts * cap * s->desc.part_rec_bytes;
        size_t img_bytes = nparts * image_bytes;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (seg_cap != 0 && rec_bytes + img_bytes + 8 * nparts >= free_b / 2)
        {
            /* not enough memory for the padded segments: global cursors */
            seg_cap = seg_max = 0;
            cap = ((size_t)((double)max_rows / nparts * 1.25) + 64 + 3) & ~(size_t)3;
            rec_bytes = nparts * cap * s->desc.part_rec_bytes;
        }
        if (nparts < 0x7fffffffULL && cap < 0x7fffffffULL &&
            rec_bytes + img_bytes + 8 * nparts < free_b / 2)
        {
            /* cursors 32 bytes apart: see PGS_PART_CURSOR */
            const char *ecs = getenv("PGSTROM_CURSOR_SHIFT");
            s->gs.part_pad = (cl_uint)(ecs ? std::min(5, std::max(0, atoi(ecs))) : 3);
            CUDA_CHECK(cudaMalloc((void **)&s->gs.part_cursor, (4 * nparts) << s->gs.part_pad));
            CUDA_CHECK(cudaMalloc((void **)&s->gs.part_nused, 4 * nparts));
            CUDA_CHECK(cudaMalloc((void **)&s->gs.part_recs, rec_bytes));
            CUDA_CHECK(cudaMalloc((void **)&s->gs.part_images, img_bytes));
            if (seg_cap != 0)
            {
                CUDA_CHECK(cudaMalloc((void **)&s->gs.part_seg_counts, 2 * nparts * seg_max));
                CUDA_CHECK(cudaMemset(s->gs.part_seg_counts, 0, 2 * nparts * seg_max));
            }
            s->gs.part_nparts = (cl_uint)nparts;
            s->gs.part_cap = (cl_uint)cap;
            s->gs.part_slots = (cl_uint)slots;
            s->gs.part_seg_cap = (cl_uint)seg_cap;
            s->gs.part_seg_max = (cl_uint)seg_max;
            s->smem_partagg = std::max<size_t>(128, s->desc.partagg_head_bytes) + image_bytes;
            CUDA_CHECK(cudaFuncSetAttribute((const void *)s->k_partagg,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)s->smem_partagg));
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                           &per_sm, (const void *)s->k_partagg, 256, s->smem_partagg));
            s->grid_partagg = (int)std::min<size_t>(nparts, (size_t)s->num_sms * std::max(1, per_sm));
        }
    }
    if (s->desc.num_keys > 0)
    {
        /* with partitions the global table only takes what they refuse.  The
         * planner's number is an estimate: the table is sized for it, and
         * for what the CTA-local tables can spill at the end of one launch,
         * and grows when the scan proves the estimate wrong
         * (session_grow_table) */
        double want = s->config.num_groups * 2.0;
        if (s->gs.part_nparts)
            want = std::max(s->config.num_groups * 0.125, std::min(want, 1048576.0));
        size_t spill = (size_t)s->grid_main * (s->sh_nslots - s->sh_nslots / 4);
        size_t nslots = 1024;
        size_t free_b = 0, total_b = 0;
        if (want < 2.0 * (double)(s->sh_nslots - s->sh_nslots / 4))
            want = 2.0 * (double)(s->sh_nslots - s->sh_nslots / 4);
        while ((double)nslots < want && nslots < (1ULL << 31))
            nslots <<= 1;
        cudaMemGetInfo(&free_b, &total_b);
        while (nslots > 1024 && nslots * (size_t)s->desc.slot_stride_bytes > free_b / 2)
            nslots >>= 1;
        CUDA_CHECK(cudaMalloc((void **)&s->gs.gh_slots, nslots * (size_t)s->desc.slot_stride_bytes));
        s->gs.gh_nslots = (cl_uint)nslots;
        s->gs.gh_max_probe = (cl_uint)std::min<size_t>(nslots, 4096);
        /* overflow log: what one launch can spill, for every chunk in flight */
        size_t ovf_cap = std::max<size_t>(65536, spill * 2);
        CUDA_CHECK(cudaMalloc((void **)&s->gs.ovf_recs, ovf_cap * (size_t)s->desc.slot_bytes));
        s->gs.ovf_cap = (cl_uint)ovf_cap;
    }
    return StromError_Success;
}

/*
 * The scan met more groups than the global table was sized for: allocate a
 * larger table (and a fresh overflow log), move the groups over with one
 * kernel, all stream ordered behind the chunks in flight - those were
 * launched with the old pointers and finish before the move starts.  The old
 * buffers are freed when the session next synchronises.  Failure to allocate
 * is not an error: rows that find no room are re-checked by the host.
 */
static int
session_grow_table(pgs_session *s, size_t min_slots)
{
    if (s->desc.num_keys == 0)
        return StromError_Success;
    size_t nslots = s->gs.gh_nslots;
    while (nslots < min_slots && nslots < (1ULL << 31))
        nslots <<= 1;
    if (nslots <= s->gs.gh_nslots)
        return StromError_Success;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t bytes = nslots * (size_t)s->desc.slot_stride_bytes;
    size_t log_bytes = (size_t)s->gs.ovf_cap * s->desc.slot_bytes;
    if (bytes + log_bytes > free_b / 2)
        return StromError_Success;
    pgs_gstate to = s->gs;
    if (cudaMalloc((void **)&to.gh_slots, bytes) != cudaSuccess ||
        cudaMalloc((void **)&to.ovf_recs, log_bytes) != cudaSuccess)
    {
        cudaGetLastError();
        if (to.gh_slots != s->gs.gh_slots)
            cudaFree(to.gh_slots);
        return StromError_Success;
    }
    to.gh_nslots = (cl_uint)nslots;
    to.gh_max_probe = (cl_uint)std::min<size_t>(nslots, 4096);
    s->ovf_which ^= 1;
    to.ovf_count = (cl_uint *)((char *)s->d_scratch + 40 + 8 * s->ovf_which);
    if (!s->d_kg_misc)
        CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
    /* an empty table: slots zeroed = EMPTY, cells get their identity from the
     * init kernel run on a state that has only this table */
    {
        pgs_gstate only = to;
        only.part_nparts = 0;
        only.part_slots = 0;
        /* (the init kernel also clears the counters it is given: hand it
         * scratch words of its own - the live ones stay) */
        only.ng_state = s->gs.ng_partial;       /* harmless scratch rows */
        only.ng_ticket = (cl_uint *)((char *)s->d_scratch + 96);
        only.gh_ngroups = (cl_uint *)((char *)s->d_scratch + 100);
        only.gh_nused = (cl_uint *)((char *)s->d_scratch + 104);
        void *args[] = { &only };
        int grid = std::max(1, std::min<int>(s->num_sms * 8, (int)((nslots + 255) / 256)));
        int rc = launch_kernel(s, s->k_init, grid, 256, 0, args);
        if (rc != StromError_Success)
            return rc;
    }
    CUDA_CHECK(cudaMemsetAsync(s->gs.gh_nused, 0, sizeof(cl_uint), s->s_exec));
    CUDA_CHECK(cudaMemsetAsync(s->d_kg_misc, 0, sizeof(kern_gpupreagg), s->s_exec));
    {
        void *args[] = { &s->gs, &to, &s->d_kg_misc };
        size_t n = (size_t)s->gs.gh_nslots + s->gs.ovf_cap;
        int grid = std::max(1, std::min<int>(s->num_sms * 8, (int)((n + 255) / 256)));
        int rc = launch_kernel(s, s->k_rehash, grid, 256, 0, args);
        if (rc != StromError_Success)
            return rc;
    }
    CUDA_CHECK(cudaMemsetAsync(s->gs.ovf_count, 0, sizeof(cl_uint), s->s_exec));
    s->graveyard.push_back(s->gs.gh_slots);
    s->graveyard.push_back(s->gs.ovf_recs);
    s->gs = to;
    s->num_table_grown++;
    return StromError_Success;
}

static void
session_bury(pgs_session *s)
{
    /* call only after the exec stream has been synchronised */
    for (void *p : s->graveyard)
        cudaFree(p);
    s->graveyard.clear();
}

static int
launch_kernel(pgs_session *s, cudaKernel_t k, int grid, int block, size_t smem,
              void **args)
{
    CUDA_CHECK(cudaLaunchKernel((const void *)k, dim3(grid), dim3(block), args,
                                smem, s->s_exec));
    s->launches++;
    return StromError_Success;
}

/* slots of the persistent GROUP BY state: global table + table images */
static size_t
state_nslots(const pgs_session *s)
{
    return (size_t)s->gs.gh_nslots + (size_t)s->gs.part_nparts * s->gs.part_slots;
}

/* The key heap of a program that groups by text / bpchar columns: values of
 * more than 7 bytes are stored once in HBM and travel as 8-byte words
 * (kern_textlib.cuh; the reference moves varlena keys as toast offsets,
 * opencl_gpupreagg.h:326-366).  pg_strom.key_heap_size (MB; 0 = no heap:
 * rows with a long key are re-checked on the host) is the least room for the
 * strings; the heap and the lookup table follow the planner's estimate of
 * the number of groups beyond that. */
static int
session_alloc_keyheap(pgs_session *s)
{
    long long mb = pgs::guc_int("pg_strom.key_heap_size");
    if (s->desc.num_text_keys == 0 || mb <= 0)
        return StromError_Success;
    void   *d_ctl = NULL;
    size_t  ctl_bytes = 0;
    if (cudaLibraryGetGlobal(&d_ctl, &ctl_bytes, s->library, "pgs_keyheap") != cudaSuccess ||
        ctl_bytes != sizeof(pgs_keyheap_ctl))
    {
        cudaGetLastError();
        return StromError_Success;      /* program without the lookup: long keys are re-checked */
    }
    size_t nslots = 1u << 16;
    double want = 4.0 * std::max(1.0, s->config.num_groups) * s->desc.num_text_keys;
    while (nslots < (1u << 26) && (double)nslots < want)
        nslots <<= 1;
    char *base = NULL;
    size_t heap_bytes = (size_t)mb << 20;
    {
        /* the GUC is the floor; a plan that expects many groups gets room for
         * 64 bytes per expected key (8-byte length + a 56-byte string), as far
         * as a quarter of the free device memory goes */
        size_t free_b = 0, total_b = 0;
        double by_estimate = 64.0 * std::max(1.0, s->config.num_groups) * s->desc.num_text_keys;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
        {
            cudaGetLastError();
            free_b = 0;
        }
        if (by_estimate > (double)heap_bytes)
            heap_bytes = (size_t)std::min(by_estimate, (double)(free_b / 4));
        heap_bytes = std::max(heap_bytes, (size_t)mb << 20) & ~(size_t)127;
    }
    if (cudaMalloc((void **)&base, 16 * nslots + 128 + heap_bytes) != cudaSuccess)
    {
        /* no room for the heap on this device: the query still runs, rows
         * with a long key go to the host */
        cudaGetLastError();
        return StromError_Success;
    }
    s->kh.slots = (cl_ulong *)base;
    s->kh.heap_used = (cl_ulong *)(base + 16 * nslots);
    s->kh.heap = (unsigned char *)(base + 16 * nslots + 128);
    s->kh.heap_bytes = heap_bytes;
    s->kh.nslots = (cl_uint)nslots;
    s->kh.max_probe = 512;
    CUDA_CHECK(cudaMemcpy(d_ctl, &s->kh, sizeof(pgs_keyheap_ctl), cudaMemcpyHostToDevice));
    return StromError_Success;
}

static int
session_init_state(pgs_session *s)
{
    if (s->kh.nslots > 0)
    {
        /* a new scan starts with an empty key heap (stream ordered, like the
         * state itself; what the host copied at the last finish stays valid) */
        CUDA_CHECK(cudaMemsetAsync(s->kh.slots, 0, 16 * (size_t)s->kh.nslots, s->s_exec));
        CUDA_CHECK(cudaMemsetAsync(s->kh.heap_used, 0, sizeof(cl_ulong), s->s_exec));
    }
    void *args[] = { &s->gs };
    int grid = std::max(1, std::min<int>(s->num_sms * 8,
                                         (int)((state_nslots(s) + 255) / 256)));
    return launch_kernel(s, s->k_init, grid, 256, 0, args);
}

extern "C" int
pgs_preagg_open(pgs_program *program, const kern_parambuf *kparams,
                const pgs_session_config *config, pgs_session **session)
{
    if (!program || !kparams || !config || !session)
    {
        set_error("pgs_preagg_open: bad arguments");
        return StromError_BadRequestMessage;
    }
    int ord = device_ordinal(config->device);
    if (ord < 0)
    {
        set_error("pgs_preagg_open: device layer is not initialised "
                  "(pgs_cuda_init) or device index %d is out of range", config->device);
        return StromError_ServerNotReady;
    }
    pgs_session *s = new pgs_session;
    s->device_index = config->device;
    s->ordinal = ord;
    s->program = program;
    s->config = *config;
    s->kparams.assign((const char *)kparams, (const char *)kparams + kparams->length);
    if (config->result_colmeta && config->result_ncols > 0)
        s->result_colmeta.assign(config->result_colmeta,
                                 config->result_colmeta + config->result_ncols);
    s->num_sms = devices[config->device].prop.multiProcessorCount;
    s->time_kern_build_ms = program->build_ms;
    s->perfmon = pgs::guc_bool("pg_strom.perfmon");

#define OPEN_CHECK(call)                                                \
    do {                                                                \
        cudaError_t __rc = (call);                                      \
        if (__rc != cudaSuccess)                                        \
        {                                                               \
            set_error("%s failed: %s (%s:%d)", #call,                   \
                      cudaGetErrorString(__rc), __FILE__, __LINE__);    \
            pgs_preagg_close(s);                                        \
            return StromError_CudaInternal;                             \
        }                                                               \
    } while (0)

    OPEN_CHECK(cudaSetDevice(ord));
    {
        /* experiment: how much DRAM an L2 miss of a gathered sector pulls in */
        const char *eg = getenv("PGSTROM_L2_FETCH_GRANULARITY");
        if (eg && atoi(eg) > 0)
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(eg));
    }
    OPEN_CHECK(cudaLibraryLoadData(&s->library, program->cubin.data(),
                                   NULL, NULL, 0, NULL, NULL, 0));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_main, s->library, "gpupreagg_main"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_rowmap, s->library, "gpupreagg_main_rowmap"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_heap, s->library, "gpupreagg_main_heap"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_partagg, s->library, "gpupreagg_partagg"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_init, s->library, "gpupreagg_init_state"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_flush, s->library, "gpupreagg_flush"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_export, s->library, "gpupreagg_export"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_import, s->library, "gpupreagg_import"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_import_blocks, s->library, "gpupreagg_import_blocks"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_describe, s->library, "gpupreagg_describe"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_heap_index, s->library, "gpupreagg_heap_index"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_heap_staged, s->library, "gpupreagg_main_heap_staged"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_rehash, s->library, "gpupreagg_rehash"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_export_parts, s->library, "gpupreagg_export_parts"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_peer_push, s->library, "gpupreagg_peer_push"));
    OPEN_CHECK(cudaLibraryGetKernel(&s->k_peer_pull, s->library, "gpupreagg_peer_pull"));
    OPEN_CHECK(cudaStreamCreateWithFlags(&s->s_copy, cudaStreamNonBlocking));
    OPEN_CHECK(cudaStreamCreateWithFlags(&s->s_exec, cudaStreamNonBlocking));
    /* the parameter buffer also goes to the program's constant memory (the
     * kernels read KPARAM_n from there, see kern_gpupreagg.cuh) */
    {
        void   *d_const = NULL;
        size_t  const_bytes = 0;
        OPEN_CHECK(cudaLibraryGetGlobal(&d_const, &const_bytes, s->library, "pgs_const_kparams"));
        if (s->kparams.size() > const_bytes)
        {
            set_error("pgs_preagg_open: parameter buffer of %zu bytes does not fit the "
                      "%zu bytes of constant memory reserved for it",
                      s->kparams.size(), const_bytes);
            pgs_preagg_close(s);
            return StromError_BadRequestMessage;
        }
        OPEN_CHECK(cudaMemcpy(d_const, s->kparams.data(), s->kparams.size(),
                              cudaMemcpyHostToDevice));
    }

    /* ask the program about its layouts */
    {
        pgs_kern_desc *d_desc;
        OPEN_CHECK(cudaMalloc((void **)&d_desc, sizeof(pgs_kern_desc)));
        void *args[] = { &d_desc };
        cudaError_t rc = cudaLaunchKernel((const void *)s->k_describe, dim3(1), dim3(1),
                                          args, 0, s->s_exec);
        if (rc == cudaSuccess)
            rc = cudaMemcpyAsync(&s->desc, d_desc, sizeof(pgs_kern_desc),
                                 cudaMemcpyDeviceToHost, s->s_exec);
        if (rc == cudaSuccess)
            rc = cudaStreamSynchronize(s->s_exec);
        cudaFree(d_desc);
        OPEN_CHECK(rc);
        s->launches++;
    }
    if ((s->desc.num_keys > 0) != (config->needs_grouping != 0))
    {
        set_error("pgs_preagg_open: needs_grouping=%d does not match the program (%u keys)",
                  config->needs_grouping, s->desc.num_keys);
        pgs_preagg_close(s);
        return StromError_BadRequestMessage;
    }
    if (config->result_ncols > 0 && (cl_uint)config->result_ncols != s->desc.num_outcols)
    {
        set_error("pgs_preagg_open: result has %d columns, the program produces %u",
                  config->result_ncols, s->desc.num_outcols);
        pgs_preagg_close(s);
        return StromError_BadRequestMessage;
    }
    /* launch shape of the main kernel: persistent CTAs (a multiple of the SM
     * count); the 227 KB of shared memory are split between the CTA-local
     * hash table (GROUP BY) and the TMA staging ring */
    {
        size_t smem_max = devices[config->device].prop.sharedMemPerBlockOptin;
        size_t head = s->desc.static_smem_bytes;
        size_t per1k = s->desc.stage_bytes;         /* staging bytes per 1024 rows */
        int per_sm = 0;
        size_t table_bytes = 0;

        s->sh_nslots = 0;
        if (s->desc.num_keys > 0)
        {
            /* the table is used up to 75%; 1.7 x groups slots (any multiple
             * of 32) keeps linear-probe chains short without eating the
             * staging ring (measured: 1408 / 1696 / 2048 slots for 1000
             * groups -> 0.66 / 0.51 / 0.84 ms per 50M rows).  A CTA-local
             * table only pays when most groups fit. */
            double want = std::max(64.0, config->num_groups * 1.7);
            /* the ring keeps at least 2 stages of 2048 rows: with smaller
             * tiles half the lanes of a 128-row step have nothing to do */
            size_t avail = (smem_max > head + 4 * per1k + 1024
                            ? smem_max - head - 4 * per1k - 1024 : 0);
            size_t maxslots = (avail / s->desc.sh_slot_bytes) & ~(size_t)31;
            size_t nslots = ((size_t)want + 31) & ~(size_t)31;
            const char *env = getenv("PGSTROM_SH_SLOTS");
            if (env)
                nslots = ((size_t)atol(env) + 31) & ~(size_t)31;
            if (maxslots == 0 || (!env && config->num_groups > 3.0 * (double)maxslots))
                nslots = 0;
            if (nslots > maxslots)
                nslots = maxslots;
            s->sh_nslots = (cl_uint)nslots;
            table_bytes = nslots * s->desc.sh_slot_bytes;
            /* very many groups, no WHERE clause: the deal pass keeps the
             * cursors of its record segments in shared memory (16 bits per
             * partition), in the place of the CTA-local table */
            s->seg_cursor_bytes = 0;
            /* (measured on B200, 100 M rows / 10 M groups: 5.78 ms against
             * 5.33 ms with global cursors - the 2.9 M segment tails no longer
             * fit L2, so the 32-byte record stores reach DRAM one by one
             * instead of as merged 128-byte lines, which costs more than the
             * cursor atomics did.  Kept behind PGSTROM_SEGMENTS=1.) */
            if (nslots == 0 && !s->desc.has_qual && getenv("PGSTROM_SEGMENTS") &&
                atoi(getenv("PGSTROM_SEGMENTS")) != 0)
            {
                size_t nparts = partition_plan(s, NULL);
                if (nparts != 0 && 2 * nparts <= 65536)
                {
                    s->seg_cursor_bytes = (4 * ((nparts + 1) / 2) + 127) & ~(size_t)127;
                    table_bytes = s->seg_cursor_bytes;
                }
            }
        }
        /* staging ring: one CTA (1 producer + 16 consumer warps) per SM owns
         * the whole shared memory; prefer 3 stages of 4096 rows (measured
         * best on B200), shrink the tile, then the depth, to fit */
        {
            size_t budget = smem_max - head - table_bytes - 1024;
            size_t target = budget;
            cl_uint tile = 4096, stages = 3;
            const char *et = getenv("PGSTROM_TILE_ROWS");
            const char *es = getenv("PGSTROM_NUM_STAGES");
            if (target > budget)
                target = budget;
            if (s->desc.max_tile_rows > tile)
            {
                /* only the qual's columns are staged (gather variant): a
                 * deeper ring of longer tiles keeps more bytes in flight */
                tile = (s->desc.max_tile_rows / 1024) * 1024;
                stages = 4;
            }
            while (stages > 2 && (size_t)stages * (tile / 1024) * per1k > target)
                stages--;
            while (tile > 1024 && (size_t)stages * (tile / 1024) * per1k > target)
                tile -= 1024;
            if (s->seg_cursor_bytes != 0)
            {
                /* one step of 128 rows per warp and tile: every warp has the
                 * same share, and shared-memory cursors need no batching */
                tile = (cl_uint)std::max<size_t>(1024, ((size_t)s->desc.block_threads - 32) * 4 / 1024 * 1024);
                stages = (cl_uint)std::min<size_t>(4, std::max<size_t>(2, target / ((tile / 1024) * per1k)));
                while (tile > 1024 && (size_t)stages * (tile / 1024) * per1k > target)
                    tile -= 1024;
            }
            if (et) tile = (cl_uint)std::max(1024, (atoi(et) / 1024) * 1024);
            if (es) stages = (cl_uint)std::min(8, std::max(1, atoi(es)));
            s->tile_rows = tile;
            s->nstages = stages;
        }
        s->smem_main = head + (size_t)s->nstages * (s->tile_rows / 1024) * per1k + table_bytes;
        if (s->smem_main > smem_max)
        {
            set_error("device program needs %zu bytes of shared memory, the device allows %zu",
                      s->smem_main, smem_max);
            pgs_preagg_close(s);
            return StromError_OutOfSharedMemory;
        }
        OPEN_CHECK(cudaFuncSetAttribute((const void *)s->k_main,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)s->smem_main));
        OPEN_CHECK(cudaFuncSetAttribute((const void *)s->k_rowmap,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)s->smem_main));
        OPEN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                       &per_sm, (const void *)s->k_main,
                       (int)s->desc.block_threads, s->smem_main));
        if (per_sm < 1)
            per_sm = 1;
        const char *env = getenv("PGSTROM_CTAS_PER_SM");
        if (env && atoi(env) > 0)
            per_sm = std::min(per_sm, atoi(env));
        s->grid_main = s->num_sms * per_sm;
        /* the heap-page kernel reads the pages straight from HBM (a chain of
         * dependent loads per tuple: row item, line pointer, header,
         * attributes): it hides that latency with resident warps, not with a
         * ring, so it only asks for the CTA-local table and the scratch of
         * the block reduction and runs as many CTAs per SM as registers allow */
        {
            size_t scratch = 8 * (size_t)(1 + s->desc.num_cells) * (s->desc.block_threads / 32 + 1);
            int heap_per_sm = 0;
            s->smem_heap = head + std::max(table_bytes, scratch) + 128;
            OPEN_CHECK(cudaFuncSetAttribute((const void *)s->k_heap,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)s->smem_heap));
            OPEN_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                           &heap_per_sm, (const void *)s->k_heap,
                           PGS_HEAP_BLOCK_THREADS, s->smem_heap));
            s->grid_heap = s->num_sms * std::max(1, heap_per_sm);
        }
        /* heap pages through the ring (KDS_FORMAT_ROW without a row map):
         * stages of 4 pages + their row items, as many as fit next to the
         * CTA-local table (at least 3, else 2-page stages, else not at all) */
        {
            const char *eh = getenv("PGSTROM_HEAP_STAGED");
            size_t budget = smem_max - head - table_bytes - 1024;
            s->heap_pps = 0;
            if (!(eh && atoi(eh) == 0))
            {
                for (cl_uint pps = 4; pps >= 2 && s->heap_pps == 0; pps -= 2)
                {
                    size_t sb = (size_t)pps * BLCKSZ +
                        (((size_t)4 * (pps * 292 + 4) + 127) & ~(size_t)127) + 128;
                    size_t n = std::min<size_t>(budget / sb, 8);
                    if (n >= 3)
                    {
                        s->heap_pps = pps;
                        s->heap_nstages = (cl_uint)n;
                        s->smem_heap_staged = head + n * sb + table_bytes;
                    }
                }
            }
            if (s->heap_pps != 0)
                OPEN_CHECK(cudaFuncSetAttribute((const void *)s->k_heap_staged,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)s->smem_heap_staged));
        }
    }
    {
        int rc = session_alloc_state(s);
        if (rc == StromError_Success)
            rc = session_alloc_keyheap(s);
        if (rc == StromError_Success)
            rc = session_init_state(s);
        if (rc != StromError_Success)
        {
            pgs_preagg_close(s);
            return rc;
        }
    }
    int nslots = config->max_async_chunks > 0 ? config->max_async_chunks
                                              : (int)pgs::guc_int("pg_strom.max_async_chunks");
    s->slots.resize(std::max(1, nslots));
    for (auto &sl : s->slots)
    {
        OPEN_CHECK(cudaEventCreateWithFlags(&sl.ev_copied, cudaEventDisableTiming));
        OPEN_CHECK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
        OPEN_CHECK(cudaEventCreate(&sl.ev_k0));
        OPEN_CHECK(cudaEventCreate(&sl.ev_k1));
        OPEN_CHECK(cudaHostAlloc((void **)&sl.h_status, 64, cudaHostAllocPortable));
    }
    OPEN_CHECK(cudaStreamSynchronize(s->s_exec));
    *session = s;
    return StromError_Success;
}

static int
slot_reserve(pgs_session *s, ChunkSlot &sl, size_t kds_bytes, size_t kg_bytes,
             uint32_t nitems, bool need_kds)
{
    if (need_kds && sl.d_kds_cap < kds_bytes)
    {
        if (sl.d_kds)
            CUDA_CHECK(cudaFree(sl.d_kds));
        sl.d_kds = NULL;
        sl.d_kds_cap = 0;
        size_t cap = std::max(kds_bytes, s->config.max_chunk_bytes);
        CUDA_CHECK(cudaMalloc(&sl.d_kds, cap));
        sl.d_kds_cap = cap;
    }
    if (sl.kg_cap < kg_bytes)
    {
        if (sl.d_kgpreagg) CUDA_CHECK(cudaFree(sl.d_kgpreagg));
        if (sl.h_kgpreagg) CUDA_CHECK(cudaFreeHost(sl.h_kgpreagg));
        sl.d_kgpreagg = NULL;
        sl.h_kgpreagg = NULL;
        size_t cap = (kg_bytes + 4095) & ~(size_t)4095;
        CUDA_CHECK(cudaMalloc((void **)&sl.d_kgpreagg, cap));
        CUDA_CHECK(cudaHostAlloc((void **)&sl.h_kgpreagg, cap, cudaHostAllocPortable));
        sl.kg_cap = cap;
    }
    size_t words = ((size_t)nitems + 31) / 32 + 1;
    if (sl.recheck_words < words)
    {
        if (sl.d_recheck) CUDA_CHECK(cudaFree(sl.d_recheck));
        sl.d_recheck = NULL;
        size_t cap = std::max(words, ((size_t)s->config.max_chunk_rows + 31) / 32 + 1);
        CUDA_CHECK(cudaMalloc((void **)&sl.d_recheck, cap * 4));
        CUDA_CHECK(cudaMemsetAsync(sl.d_recheck, 0, cap * 4, s->s_exec));
        sl.recheck_words = cap;
    }
    return StromError_Success;
}

/* completion of whatever occupies the slot (the reference's
 * clserv_respond_gpupreagg, gpupreagg.c:3009-3194) */
static int
slot_retire(pgs_session *s, ChunkSlot &sl)
{
    if (!sl.busy)
        return StromError_Success;
    CUDA_CHECK(cudaEventSynchronize(sl.ev_done));
    ChunkResult &res = s->results[sl.ticket];
    res.status = *sl.h_status;
    res.done = true;
    if (s->desc.num_keys > 0 && sl.table_epoch == s->num_table_grown)
    {
        /* table occupancy and overflow log as the chunk left them (copied
         * behind its status word): a table more than half full, or one that
         * already turned states away, is replaced by a larger one before the
         * next chunk is launched */
        const uint32_t *w = (const uint32_t *)sl.h_status;
        size_t nused = w[2];
        size_t nlog = std::max(w[4], w[6]);
        if (nlog > 0 || nused * 2 > (size_t)s->gs.gh_nslots)
        {
            int rc = session_grow_table(s, 4 * std::max<size_t>(nused + nlog, s->gs.gh_nslots));
            if (rc != StromError_Success)
                return rc;
        }
    }
    if (sl.timed)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sl.ev_k0, sl.ev_k1) == cudaSuccess)
        {
            s->time_kern_main_ms += ms;
            s->num_kern_main++;
            s->rows_kern_main += sl.nitems;
        }
        sl.timed = false;
    }
    if (res.status == StromError_CpuReCheck)
    {
        /* pull the per-row re-check bitmap and clear it for the next use */
        size_t words = ((size_t)sl.nitems + 31) / 32;
        std::vector<uint32_t> bm(words);
        CUDA_CHECK(cudaMemcpyAsync(bm.data(), sl.d_recheck, words * 4,
                                   cudaMemcpyDeviceToHost, s->s_exec));
        CUDA_CHECK(cudaMemsetAsync(sl.d_recheck, 0, words * 4, s->s_exec));
        CUDA_CHECK(cudaStreamSynchronize(s->s_exec));
        s->num_dma_recv++;
        s->bytes_dma_recv += words * 4;
        for (size_t w = 0; w < words; w++)
        {
            uint32_t bits = bm[w];
            while (bits)
            {
                int b = __builtin_ctz(bits);
                res.recheck_rows.push_back((uint32_t)(w * 32 + b));
                bits &= bits - 1;
            }
        }
        s->num_rechecked_chunks++;
    }
    sl.busy = false;
    return StromError_Success;
}

static int
submit_common(pgs_session *s, const kern_data_store *kds_host, const void *kds_dev,
              size_t length, uint32_t nitems, int format, const kern_row_map *krowmap,
              pgs_ticket *ticket)
{
    if (!s || s->aborted)
    {
        set_error("session is closed or aborted");
        return StromError_BadRequestMessage;
    }
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    s->push_pending = false;    /* rows after a push: the state is not empty any more */
    pgs_ticket t = s->next_ticket;
    ChunkSlot &sl = s->slots[(size_t)(t % (pgs_ticket)s->slots.size())];
    int rc = slot_retire(s, sl);
    if (rc != StromError_Success)
        return rc;

    size_t rowmap_bytes = sizeof(cl_int);
    if (krowmap && krowmap->nvalids >= 0)
        rowmap_bytes += sizeof(cl_int) * (size_t)krowmap->nvalids;
    size_t kg_head = STROMALIGN(offsetof(kern_gpupreagg, kparams) + s->kparams.size());
    size_t kg_bytes = kg_head + rowmap_bytes;

    rc = slot_reserve(s, sl, length, kg_bytes, nitems, kds_host != NULL);
    if (rc != StromError_Success)
        return rc;
    /* kern_gpupreagg image: status, kparams, row map */
    kern_gpupreagg *kg = (kern_gpupreagg *)sl.h_kgpreagg;
    memset(kg, 0, offsetof(kern_gpupreagg, kparams));
    kg->status = StromError_Success;
    memcpy(&kg->kparams, s->kparams.data(), s->kparams.size());
    kern_row_map *rm = (kern_row_map *)(sl.h_kgpreagg + kg_head);
    if (krowmap && krowmap->nvalids >= 0)
    {
        rm->nvalids = krowmap->nvalids;
        memcpy(rm->rindex, krowmap->rindex, sizeof(cl_int) * (size_t)krowmap->nvalids);
    }
    else
        rm->nvalids = -1;

    const void *d_kds = kds_dev;
    if (kds_host)
    {
        /* host chunk: H2D on the copy stream, so that it overlaps the kernel
         * of the previous chunk; the exec stream waits for it */
        if (format == KDS_FORMAT_ROW)
        {
            /* clserv_dmasend_data_store (datastore.c:908-968): head, block
             * items and row items in one piece, then the heap pages - which
             * live wherever bitem->page says (shared buffers) - behind the
             * next BLCKSZ boundary; runs of adjacent pages go as one copy */
            size_t head_len = (size_t)((const char *)KERN_DATA_STORE_ROWITEM(kds_host, kds_host->nitems) -
                                       (const char *)kds_host);
            size_t offset = (size_t)((const char *)KERN_DATA_STORE_ROWBLOCK(kds_host, 0) -
                                     (const char *)kds_host);
            const kern_blkitem *bitem = KERN_DATA_STORE_BLKITEM(kds_host, 0);

            CUDA_CHECK(cudaMemcpyAsync(sl.d_kds, kds_host, head_len,
                                       cudaMemcpyHostToDevice, s->s_copy));
            s->num_dma_send++;
            s->bytes_dma_send += head_len;
            for (cl_uint i = 0, n = 0; i < kds_host->nblocks; i++)
            {
                if (i + 1 < kds_host->nblocks &&
                    bitem[i].page + BLCKSZ == bitem[i + 1].page)
                {
                    n++;
                    continue;
                }
                CUDA_CHECK(cudaMemcpyAsync((char *)sl.d_kds + offset,
                                           (const void *)(uintptr_t)bitem[i - n].page,
                                           (size_t)BLCKSZ * (n + 1),
                                           cudaMemcpyHostToDevice, s->s_copy));
                s->num_dma_send++;
                s->bytes_dma_send += (size_t)BLCKSZ * (n + 1);
                offset += (size_t)BLCKSZ * (n + 1);
                n = 0;
            }
        }
        else
        {
            CUDA_CHECK(cudaMemcpyAsync(sl.d_kds, kds_host, length,
                                       cudaMemcpyHostToDevice, s->s_copy));
            s->num_dma_send++;
            s->bytes_dma_send += length;
        }
        d_kds = sl.d_kds;
        CUDA_CHECK(cudaMemcpyAsync(sl.d_kgpreagg, sl.h_kgpreagg, kg_bytes,
                                   cudaMemcpyHostToDevice, s->s_copy));
        CUDA_CHECK(cudaEventRecord(sl.ev_copied, s->s_copy));
        CUDA_CHECK(cudaStreamWaitEvent(s->s_exec, sl.ev_copied, 0));
    }
    else
    {
        /* chunk already in HBM: nothing to overlap, keep it on one stream */
        CUDA_CHECK(cudaMemcpyAsync(sl.d_kgpreagg, sl.h_kgpreagg, kg_bytes,
                                   cudaMemcpyHostToDevice, s->s_exec));
    }
    s->num_dma_send++;
    s->bytes_dma_send += kg_bytes;

    bool use_rowmap = (krowmap && krowmap->nvalids >= 0);
    bool use_heap = (format == KDS_FORMAT_ROW || format == KDS_FORMAT_ROW_FLAT);
    /* segment mode belongs to gpupreagg_main alone (its grid is what the
     * record areas are cut for); every other scan kernel deals through the
     * global cursors */
    const bool seg_launch = (s->gs.part_seg_cap != 0 && !use_rowmap && !use_heap);
    pgs_gstate gs_launch = s->gs;
    if (!seg_launch)
        gs_launch.part_seg_cap = 0;
    void *args[] = { &sl.d_kgpreagg, (void *)&d_kds, &gs_launch, &sl.d_recheck, &s->sh_nslots,
                     &s->tile_rows, &s->nstages };
    int grid = s->grid_main;
    cl_uint no_stages = 0;
    if (use_heap)
    {
        /* one thread per tuple, grid-stride */
        uint32_t nblk = (nitems + PGS_HEAP_BLOCK_THREADS - 1) / PGS_HEAP_BLOCK_THREADS;
        grid = s->grid_heap;
        if ((uint32_t)grid > nblk)
            grid = (int)std::max<uint32_t>(1, nblk);
        args[6] = &no_stages;
    }
    else if (!use_rowmap)
    {
        uint32_t ntiles = (nitems + s->tile_rows - 1) / s->tile_rows;
        if ((uint32_t)grid > ntiles)
            grid = (int)std::max<uint32_t>(1, ntiles);
    }
    if (s->perfmon)
        CUDA_CHECK(cudaEventRecord(sl.ev_k0, s->s_exec));
    if (format == KDS_FORMAT_ROW && !use_rowmap && s->heap_pps != 0)
    {
        /* pages through the staging ring: index of the first row item of
         * every page, then the staged kernel (kern_gpupreagg.cuh) */
        const size_t index_words = 65538;
        if (!sl.d_heap_index)
            CUDA_CHECK(cudaMalloc((void **)&sl.d_heap_index, index_words * 4));
        CUDA_CHECK(cudaMemsetAsync(sl.d_heap_index, 0xff, (index_words - 1) * 4, s->s_exec));
        CUDA_CHECK(cudaMemsetAsync(sl.d_heap_index + (index_words - 1), 0, 4, s->s_exec));
        {
            void *iargs[] = { (void *)&d_kds, &sl.d_heap_index };
            int igrid = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)s->num_sms * 8,
                                                                      (nitems + 255) / 256));
            rc = launch_kernel(s, s->k_heap_index, igrid, 256, 0, iargs);
            if (rc != StromError_Success)
                return rc;
        }
        void *hargs[] = { &sl.d_kgpreagg, (void *)&d_kds, &gs_launch, &sl.d_recheck, &s->sh_nslots,
                          &s->heap_pps, &s->heap_nstages, &sl.d_heap_index };
        rc = launch_kernel(s, s->k_heap_staged, s->num_sms, (int)s->desc.block_threads,
                           s->smem_heap_staged, hargs);
    }
    else
        rc = launch_kernel(s, use_heap ? s->k_heap : (use_rowmap ? s->k_rowmap : s->k_main), grid,
                           use_heap ? PGS_HEAP_BLOCK_THREADS : (int)s->desc.block_threads,
                           use_heap ? s->smem_heap : s->smem_main, args);
    if (rc != StromError_Success)
        return rc;
    if (s->gs.part_nparts != 0)
    {
        /* second pass of the partitioned GROUP BY: every partition that
         * received records is aggregated into its table image */
        cl_uint nseg = (seg_launch ? (cl_uint)grid : 0U);
        void *pargs[] = { &sl.d_kgpreagg, (void *)&d_kds, &gs_launch, &sl.d_recheck, &nseg };
        rc = launch_kernel(s, s->k_partagg, s->grid_partagg, 256, s->smem_partagg, pargs);
        if (rc != StromError_Success)
            return rc;
    }
    if (s->perfmon)
    {
        CUDA_CHECK(cudaEventRecord(sl.ev_k1, s->s_exec));
        sl.timed = true;
    }
    CUDA_CHECK(cudaMemcpyAsync(sl.h_status, sl.d_kgpreagg, sizeof(int32_t),
                               cudaMemcpyDeviceToHost, s->s_exec));
    if (s->desc.num_keys > 0)
    {
        CUDA_CHECK(cudaMemcpyAsync(sl.h_status + 2, (char *)s->d_scratch + 32, 24,
                                   cudaMemcpyDeviceToHost, s->s_exec));
        sl.table_epoch = s->num_table_grown;
    }
    CUDA_CHECK(cudaEventRecord(sl.ev_done, s->s_exec));
    s->num_dma_recv++;
    s->bytes_dma_recv += sizeof(int32_t);
    s->num_chunks++;
    sl.busy = true;
    sl.ticket = t;
    sl.nitems = nitems;
    s->results[t] = ChunkResult();
    s->next_ticket++;
    if (ticket)
        *ticket = t;
    return StromError_Success;
}

extern "C" int
pgs_preagg_submit(pgs_session *session, const kern_data_store *kds_in,
                  const kern_row_map *krowmap, pgs_ticket *ticket)
{
    if (!kds_in)
    {
        set_error("pgs_preagg_submit: no chunk");
        return StromError_BadRequestMessage;
    }
    size_t length = kds_in->length;
    if (kds_in->format == KDS_FORMAT_ROW)
    {
        /* what clserv_dmasend_data_store sends (datastore.c:908-968): head,
         * block and row items, then the pages from the next BLCKSZ boundary */
        length = (size_t)((const char *)KERN_DATA_STORE_ROWBLOCK(kds_in, kds_in->nblocks) -
                          (const char *)kds_in);
    }
    else if (kds_in->format != KDS_FORMAT_ROW_FLAT && kds_in->format != KDS_FORMAT_COLUMN)
    {
        set_error("pgs_preagg_submit: chunk format %d is not an input format "
                  "(KDS_FORMAT_ROW, ROW_FLAT or COLUMN)", (int)kds_in->format);
        return StromError_BadRequestMessage;
    }
    return submit_common(session, kds_in, NULL, length, kds_in->nitems,
                         kds_in->format, krowmap, ticket);
}

extern "C" int
pgs_preagg_submit_device(pgs_session *session, const void *kds_in_device,
                         size_t length, uint32_t nitems,
                         const kern_row_map *krowmap, pgs_ticket *ticket)
{
    return pgs_preagg_submit_device_format(session, kds_in_device, length, nitems,
                                           KDS_FORMAT_COLUMN, krowmap, ticket);
}

extern "C" int
pgs_preagg_submit_device_format(pgs_session *session, const void *kds_in_device,
                                size_t length, uint32_t nitems, int format,
                                const kern_row_map *krowmap, pgs_ticket *ticket)
{
    if (!kds_in_device)
    {
        set_error("pgs_preagg_submit_device: no chunk");
        return StromError_BadRequestMessage;
    }
    if (format != KDS_FORMAT_ROW && format != KDS_FORMAT_ROW_FLAT && format != KDS_FORMAT_COLUMN)
    {
        set_error("pgs_preagg_submit_device: chunk format %d is not an input format", format);
        return StromError_BadRequestMessage;
    }
    return submit_common(session, NULL, kds_in_device, length, nitems, format, krowmap, ticket);
}

extern "C" int
pgs_preagg_wait(pgs_session *s, pgs_ticket ticket, int timeout_ms, int32_t *status)
{
    if (!s)
        return StromError_BadRequestMessage;
    auto it = s->results.find(ticket);
    if (it == s->results.end())
    {
        set_error("unknown ticket %lld", (long long)ticket);
        return StromError_BadRequestMessage;
    }
    if (!it->second.done)
    {
        ChunkSlot &sl = s->slots[(size_t)(ticket % (pgs_ticket)s->slots.size())];
        if (sl.busy && sl.ticket == ticket)
        {
            if (timeout_ms >= 0)
            {
                /* mqueue.c:257-324: poll up to the timeout */
                auto t0 = std::chrono::steady_clock::now();
                for (;;)
                {
                    cudaError_t q = cudaEventQuery(sl.ev_done);
                    if (q == cudaSuccess)
                        break;
                    if (q != cudaErrorNotReady)
                    {
                        set_error("cudaEventQuery: %s", cudaGetErrorString(q));
                        return StromError_CudaInternal;
                    }
                    double ms = std::chrono::duration<double, std::milli>(
                        std::chrono::steady_clock::now() - t0).count();
                    if (ms >= (double)timeout_ms)
                        return -1;      /* still running */
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                }
            }
            int rc = slot_retire(s, sl);
            if (rc != StromError_Success)
                return rc;
        }
    }
    if (status)
        *status = it->second.status;
    return StromError_Success;
}

extern "C" int64_t
pgs_preagg_recheck_rows(pgs_session *s, pgs_ticket ticket, uint32_t *rows, int64_t max_rows)
{
    auto it = s->results.find(ticket);
    if (it == s->results.end() || !it->second.done)
        return -1;
    int64_t n = (int64_t)it->second.recheck_rows.size();
    if (rows)
        for (int64_t i = 0; i < n && i < max_rows; i++)
            rows[i] = it->second.recheck_rows[(size_t)i];
    return n;
}

static int
drain(pgs_session *s)
{
    for (auto &sl : s->slots)
    {
        int rc = slot_retire(s, sl);
        if (rc != StromError_Success)
            return rc;
    }
    return StromError_Success;
}

extern "C" int
pgs_preagg_finish(pgs_session *s, kern_data_store *kds_dst, int reset,
                  uint32_t *nrows_needed, int32_t *status)
{
    if (!s || !kds_dst)
        return StromError_BadRequestMessage;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    int rc = StromError_Success;
    if (kds_dst->format != KDS_FORMAT_TUPSLOT || kds_dst->ncols != s->desc.num_outcols)
    {
        set_error("pgs_preagg_finish: destination must be a TUPSLOT store with %u columns",
                  s->desc.num_outcols);
        return StromError_BadRequestMessage;
    }
    size_t head = KERN_DATA_STORE_HEAD_LENGTH(kds_dst->ncols);
    size_t stride = KERN_DATA_STORE_SLOT_STRIDE(kds_dst->ncols);
    size_t total = head + stride * (size_t)kds_dst->nrooms;
    char *d_dst = NULL;
    kern_gpupreagg *d_kg = NULL;
    int32_t h_status = 0;
    kern_data_store hdr;
    /* pinned staging: [status word | head + the first rows of the result].
     * Small results (no GROUP BY, a few thousand groups) come back with one
     * copy and one synchronisation; larger ones need a second copy once the
     * row count is known. */
    const size_t STAGE_BYTES = 256 * 1024;
    size_t first = std::min(total, STAGE_BYTES - 64);

    memset(&hdr, 0, sizeof(hdr));
    if (s->push_pending)
    {
        /* this rank pushed its state to the root (pgs_preagg_merge_peer).
         * When everything went over - always for a no-group state, for a
         * GROUP BY state when it fitted the root's exchange area - there is
         * nothing left to flush and the push kernel has reset the state: the
         * rank only waits for its own chunks and returns an empty result.
         * That keeps the ranks ahead of the root, which then never waits for
         * a straggler's flush + read-back (measured at 8 GPUs: the root's
         * merge kernel spent 0.07 ms per scan waiting) */
        s->push_pending = false;
        /* this rank is ahead of the root and its push kernel may sit waiting
         * for the root's `done` word.  Poll and yield: as quick as spinning
         * inside cudaStreamSynchronize() when the core is free, but a host
         * core that is shared with another rank's thread (the root's turn-
         * around is on the root's critical path) goes to that thread */
        cudaError_t pe;
        while ((pe = cudaStreamQuery(s->s_exec)) == cudaErrorNotReady)
            sched_yield();
        cudaGetLastError();         /* (cudaErrorNotReady is not an error) */
        if (pe != cudaSuccess)
        {
            set_error("pgs_preagg_finish: %s", cudaGetErrorString(pe));
            return StromError_CudaInternal;
        }
        if (*((volatile cl_uint *)s->h_moved) == 1)
        {
            session_bury(s);
            rc = drain(s);
            if (rc != StromError_Success)
                return rc;
            kds_dst->nitems = 0;
            if (nrows_needed)
                *nrows_needed = 0;
            if (status)
                *status = StromError_Success;
            return StromError_Success;
        }
    }
    /* the result store and the status word live as long as the session */
    if (s->d_result_cap < total)
    {
        if (s->d_result)
            CUDA_CHECK(cudaFree(s->d_result));
        s->d_result = NULL;
        s->d_result_cap = 0;
        CUDA_CHECK(cudaMalloc((void **)&s->d_result, total));
        s->d_result_cap = total;
    }
    if (!s->d_kg_misc)
        CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
    if (!s->h_result_head)
        CUDA_CHECK(cudaHostAlloc((void **)&s->h_result_head, STAGE_BYTES, cudaHostAllocPortable));
    d_dst = s->d_result;
    d_kg = s->d_kg_misc;
    kds_dst->nitems = 0;
    memcpy(s->h_result_head + 64, kds_dst, head);
    cudaError_t e = cudaMemcpyAsync(d_dst, s->h_result_head + 64, head,
                                    cudaMemcpyHostToDevice, s->s_exec);
    /* (a merge kernel before this flush left its status in the same word:
     * the first error sticks) */
    if (e == cudaSuccess && !s->merge_pending)
        e = cudaMemsetAsync(d_kg, 0, sizeof(kern_gpupreagg), s->s_exec);
    s->merge_pending = false;
    if (e == cudaSuccess)
    {
        void *args[] = { &s->gs, &d_dst, &d_kg };
        int grid = (s->desc.num_keys == 0) ? 1 :
            std::max(1, std::min<int>(s->num_sms * 8, (int)std::min<size_t>((state_nslots(s) + 255) / 256, 1u << 30)));
        rc = launch_kernel(s, s->k_flush, grid, 256, 0, args);
        if (rc != StromError_Success)
            e = cudaErrorUnknown;
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->h_result_head, d_kg, sizeof(int32_t),
                            cudaMemcpyDeviceToHost, s->s_exec);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->h_result_head + 64, d_dst, first,
                            cudaMemcpyDeviceToHost, s->s_exec);
    /* one wait for everything: the chunks still in flight, the flush and
     * the copies are ordered on the exec stream */
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s->s_exec);
    if (e == cudaSuccess)
    {
        session_bury(s);
        rc = drain(s);          /* chunk results; nothing left to wait for */
        if (rc != StromError_Success)
            return rc;
    }
    if (e == cudaSuccess)
    {
        h_status = *((int32_t *)s->h_result_head);
        memcpy(&hdr, s->h_result_head + 64, offsetof(kern_data_store, colmeta));
        if (nrows_needed)
            *nrows_needed = hdr.nitems;
        if (h_status == StromError_Success && hdr.nitems <= kds_dst->nrooms)
        {
            size_t nbytes = stride * (size_t)hdr.nitems;
            size_t have = std::min(nbytes, first - head);
            memcpy((char *)kds_dst + head, s->h_result_head + 64 + head, have);
            if (nbytes > have)
                e = cudaMemcpy((char *)kds_dst + head + have, d_dst + head + have,
                               nbytes - have, cudaMemcpyDeviceToHost);
            kds_dst->nitems = hdr.nitems;
            s->num_dma_recv += 2;
            s->bytes_dma_recv += nbytes + sizeof(hdr);
        }
        else if (h_status == StromError_Success)
            h_status = StromError_DataStoreNoSpace;
    }
    if (e == cudaSuccess && s->kh.nslots > 0 && h_status == StromError_Success)
    {
        /* the strings behind the key-heap words of the result rows
         * (pgs_preagg_key_heap / pgstrom_fixup_kernel_text_heap) */
        cl_ulong used = 0;
        e = cudaMemcpy(&used, s->kh.heap_used, sizeof(used), cudaMemcpyDeviceToHost);
        used = std::min<cl_ulong>(used, s->kh.heap_bytes);
        s->kh_host.resize((size_t)used);
        if (e == cudaSuccess && used > 0)
            e = cudaMemcpy(s->kh_host.data(), s->kh.heap, (size_t)used, cudaMemcpyDeviceToHost);
        s->kh_finishes++;
        s->num_dma_recv += 1;
        s->bytes_dma_recv += used;
    }
    if (e != cudaSuccess)
    {
        set_error("pgs_preagg_finish: %s", cudaGetErrorString(e));
        return StromError_CudaInternal;
    }
    if (status)
        *status = h_status;
    if (h_status == StromError_DataStoreNoSpace)
    {
        set_error("result store has %u rooms, %u rows are needed",
                  kds_dst->nrooms, hdr.nitems);
        return StromError_DataStoreNoSpace;
    }
    if (reset && h_status == StromError_Success)
    {
        /* stream ordered: later chunks queue up behind it.  A flush that
         * failed (result store too small) keeps the state for a retry. */
        rc = session_init_state(s);
        if (rc != StromError_Success)
            return rc;
    }
    return StromErrorIsSignificant(h_status) ? h_status : StromError_Success;
}

/* States whose text keys went through the key heap carry session-local
 * words: they cannot be merged with another session's state.  Every device
 * flushes its own partial rows instead, and PostgreSQL's final Agg merges
 * rows of one key as it does in the reference (gpupreagg.c:2169-2186). */
static int
refuse_keyheap(const pgs_session *s, const char *what)
{
    if (s && s->kh.nslots > 0)
    {
        set_error("%s: the state groups by text keys of the session's key heap and "
                  "cannot be merged across sessions; flush every device's own "
                  "partial rows (or set pg_strom.key_heap_size = 0)", what);
        return StromError_BadRequestMessage;
    }
    return StromError_Success;
}

extern "C" int
pgs_preagg_state_reset(pgs_session *s)
{
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    int rc = drain(s);
    if (rc == StromError_Success)
        rc = session_init_state(s);
    if (rc != StromError_Success)
        return rc;
    CUDA_CHECK(cudaStreamSynchronize(s->s_exec));
    return StromError_Success;
}

extern "C" int
pgs_preagg_state_export(pgs_session *s, void *device_buf, size_t buflen,
                        uint32_t *nrecords, size_t *record_bytes)
{
    if (refuse_keyheap(s, "pgs_preagg_state_export") != StromError_Success)
        return StromError_BadRequestMessage;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    int rc = drain(s);
    if (rc != StromError_Success)
        return rc;
    size_t recb = s->desc.slot_bytes;
    if (record_bytes)
        *record_bytes = recb;
    cl_uint *d_n = (cl_uint *)((char *)s->d_scratch + 64);
    cl_uint max_records = (cl_uint)std::min<size_t>(buflen / recb, 0xffffffffU);
    cl_ulong *recs = (cl_ulong *)device_buf;
    CUDA_CHECK(cudaMemsetAsync(d_n, 0, sizeof(cl_uint), s->s_exec));
    void *args[] = { &s->gs, &recs, &d_n, &max_records };
    int grid = (s->desc.num_keys == 0) ? 1 :
        std::max(1, std::min<int>(s->num_sms * 8, (int)std::min<size_t>((state_nslots(s) + 255) / 256, 1u << 30)));
    rc = launch_kernel(s, s->k_export, grid, 256, 0, args);
    if (rc != StromError_Success)
        return rc;
    cl_uint n = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n, d_n, sizeof(cl_uint), cudaMemcpyDeviceToHost, s->s_exec));
    CUDA_CHECK(cudaStreamSynchronize(s->s_exec));
    if (nrecords)
        *nrecords = n;
    if (n > max_records)
    {
        set_error("state export needs room for %u records, buffer holds %u", n, max_records);
        return StromError_DataStoreNoSpace;
    }
    return StromError_Success;
}

extern "C" int
pgs_preagg_state_import(pgs_session *s, const void *device_buf, uint32_t nrecords)
{
    if (refuse_keyheap(s, "pgs_preagg_state_import") != StromError_Success)
        return StromError_BadRequestMessage;
    if (s)
        s->push_pending = false;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    kern_gpupreagg *d_kg = NULL;
    int32_t h_status = 0;
    const cl_ulong *recs = (const cl_ulong *)device_buf;
    if (!s->d_kg_misc)
        CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
    d_kg = s->d_kg_misc;
    if (s->desc.num_keys > 0)
    {
        /* every record may be a group the table does not hold yet */
        int grc = session_grow_table(s, (size_t)s->gs.gh_nslots + 2 * (size_t)nrecords);
        if (grc != StromError_Success)
            return grc;
    }
    cudaError_t e = cudaMemsetAsync(d_kg, 0, sizeof(kern_gpupreagg), s->s_exec);
    int rc = StromError_Success;
    if (e == cudaSuccess)
    {
        void *args[] = { &s->gs, &recs, &nrecords, &d_kg };
        int grid = (s->desc.num_keys == 0) ? 1 :
            std::max(1, std::min<int>(s->num_sms * 8, (int)((nrecords + 255) / 256)));
        rc = launch_kernel(s, s->k_import, grid, 256, 0, args);
    }
    if (e == cudaSuccess && rc == StromError_Success)
        e = cudaMemcpyAsync(&h_status, d_kg, sizeof(int32_t), cudaMemcpyDeviceToHost, s->s_exec);
    if (e == cudaSuccess && rc == StromError_Success)
        e = cudaStreamSynchronize(s->s_exec);
    if (rc != StromError_Success)
        return rc;
    if (e != cudaSuccess)
    {
        set_error("pgs_preagg_state_import: %s", cudaGetErrorString(e));
        return StromError_CudaInternal;
    }
    return h_status;
}

/* ---- NCCL merge: libnccl is opened lazily so that the library loads (and
 * the single-GPU path runs) on hosts without it ---- */
struct pgs_nccl_uid { char internal[128]; };    /* == ncclUniqueId */
typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_send_fn)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_recv_fn)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_group_fn)(void);
typedef const char *(*nccl_errstr_fn)(int);
typedef int (*nccl_uid_fn)(void *);
typedef int (*nccl_init_rank_fn)(void **, int, /* ncclUniqueId by value */ struct pgs_nccl_uid, int);
typedef int (*nccl_destroy_fn)(void *);
static struct {
    void *handle;
    nccl_allgather_fn allgather;
    nccl_send_fn send;
    nccl_recv_fn recv;
    nccl_group_fn group_start, group_end;
    nccl_errstr_fn errstr;
    nccl_uid_fn get_unique_id;
    nccl_init_rank_fn comm_init_rank;
    nccl_destroy_fn comm_destroy;
} nccl;

static int
nccl_load(void)
{
    if (nccl.handle)
        return StromError_Success;
    const char *names[] = { "libnccl.so.2", "libnccl.so", NULL };
    for (int i = 0; names[i] && !nccl.handle; i++)
        nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!nccl.handle)
    {
        set_error("cannot load libnccl: %s", dlerror());
        return StromError_ServerNotReady;
    }
    nccl.allgather = (nccl_allgather_fn)dlsym(nccl.handle, "ncclAllGather");
    nccl.send = (nccl_send_fn)dlsym(nccl.handle, "ncclSend");
    nccl.recv = (nccl_recv_fn)dlsym(nccl.handle, "ncclRecv");
    nccl.group_start = (nccl_group_fn)dlsym(nccl.handle, "ncclGroupStart");
    nccl.group_end = (nccl_group_fn)dlsym(nccl.handle, "ncclGroupEnd");
    nccl.errstr = (nccl_errstr_fn)dlsym(nccl.handle, "ncclGetErrorString");
    nccl.get_unique_id = (nccl_uid_fn)dlsym(nccl.handle, "ncclGetUniqueId");
    nccl.comm_init_rank = (nccl_init_rank_fn)dlsym(nccl.handle, "ncclCommInitRank");
    nccl.comm_destroy = (nccl_destroy_fn)dlsym(nccl.handle, "ncclCommDestroy");
    if (!nccl.allgather || !nccl.send || !nccl.recv || !nccl.group_start || !nccl.group_end)
    {
        set_error("libnccl lacks a required symbol");
        return StromError_ServerNotReady;
    }
    return StromError_Success;
}

#define NCCL_CHECK(call)                                                \
    do {                                                                \
        int __rc = (call);                                              \
        if (__rc != 0)                                                  \
        {                                                               \
            set_error("%s failed: %s", #call,                           \
                      nccl.errstr ? nccl.errstr(__rc) : "?");           \
            return StromError_CudaInternal;                             \
        }                                                               \
    } while (0)

extern "C" int
pgs_nccl_get_unique_id(void *unique_id_128)
{
    int rc = nccl_load();
    if (rc != StromError_Success)
        return rc;
    if (!nccl.get_unique_id)
        return StromError_ServerNotReady;
    NCCL_CHECK(nccl.get_unique_id(unique_id_128));
    return StromError_Success;
}

extern "C" int
pgs_nccl_comm_init_rank(int device, int nranks, const void *unique_id_128, int rank,
                        void **comm)
{
    int rc = nccl_load();
    if (rc != StromError_Success)
        return rc;
    int ord = device_ordinal(device);
    if (ord < 0 || !nccl.comm_init_rank)
        return StromError_ServerNotReady;
    CUDA_CHECK(cudaSetDevice(ord));
    struct pgs_nccl_uid uid;
    memcpy(&uid, unique_id_128, sizeof(uid));
    NCCL_CHECK(nccl.comm_init_rank(comm, nranks, uid, rank));
    return StromError_Success;
}

extern "C" void
pgs_nccl_comm_destroy(void *comm)
{
    if (nccl.handle && nccl.comm_destroy && comm)
        nccl.comm_destroy(comm);
}

extern "C" int pgs_preagg_merge_exchange(pgs_session *s, void *nccl_comm, int rank, int nranks);

extern "C" int
pgs_preagg_merge_nccl(pgs_session *s, void *nccl_comm, int rank, int nranks, int root)
{
    if (refuse_keyheap(s, "pgs_preagg_merge_nccl") != StromError_Success)
        return StromError_BadRequestMessage;
    /* ncclUint8 = 1 */
    const int NCCL_UINT8 = 1;
    int rc = nccl_load();
    if (rc != StromError_Success)
        return rc;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    size_t recb = s->desc.slot_bytes;
    /*
     * Small states (no GROUP BY: one record; GROUP BY with a table of at most
     * 64K slots): every rank exports into a fixed-size block [header | cap
     * records], ONE ncclAllGather moves the blocks and the root imports them.
     * Everything is stream ordered - no allocation, no host synchronisation;
     * the caller's pgs_preagg_finish() waits once for all of it.
     */
    size_t cap = (s->desc.num_keys == 0 ? 1 : state_nslots(s));
    if (cap <= 65536 && (cap + 1) * recb * (size_t)(nranks + 1) <= ((size_t)256 << 20))
    {
        size_t block = (cap + 1) * recb;
        size_t need = block * (size_t)(nranks + 1);
        if (s->d_xchg_cap < need)
        {
            if (s->d_xchg)
                CUDA_CHECK(cudaFree(s->d_xchg));
            s->d_xchg = NULL;
            s->d_xchg_cap = 0;
            CUDA_CHECK(cudaMalloc((void **)&s->d_xchg, need));
            s->d_xchg_cap = need;
        }
        char *d_send = s->d_xchg;
        char *d_recv = s->d_xchg + block;
        cl_ulong *recs = (cl_ulong *)(d_send + recb);
        cl_uint *d_n = (cl_uint *)d_send;           /* header word = record count */
        cl_uint max_records = (cl_uint)cap;
        CUDA_CHECK(cudaMemsetAsync(d_send, 0, recb, s->s_exec));
        {
            void *args[] = { &s->gs, &recs, &d_n, &max_records };
            int grid = (s->desc.num_keys == 0) ? 1 :
                std::max(1, std::min<int>(s->num_sms * 8, (int)std::min<size_t>((state_nslots(s) + 255) / 256, 1u << 30)));
            rc = launch_kernel(s, s->k_export, grid, 256, 0, args);
            if (rc != StromError_Success)
                return rc;
        }
        NCCL_CHECK(nccl.allgather(d_send, d_recv, block, NCCL_UINT8, nccl_comm, s->s_exec));
        if (rank != root)
            return session_init_state(s);   /* this rank's state now lives on the root */
        if (!s->d_kg_misc)
            CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
        CUDA_CHECK(cudaMemsetAsync(s->d_kg_misc, 0, sizeof(kern_gpupreagg), s->s_exec));
        {
            const cl_ulong *blocks = (const cl_ulong *)d_recv;
            cl_uint nr = (cl_uint)nranks, rt = (cl_uint)root, cp = (cl_uint)cap;
            void *args[] = { &s->gs, &blocks, &nr, &rt, &cp, &s->d_kg_misc };
            int grid = (s->desc.num_keys == 0) ? 1 :
                std::max(1, std::min<int>(s->num_sms * 4, (int)((cap + 255) / 256)));
            rc = launch_kernel(s, s->k_import_blocks, grid, 256, 0, args);
            if (rc != StromError_Success)
                return rc;
        }
        /* a full root table shows up as the status of the flush that follows:
         * check it here only when the caller asked for strict merges */
        return StromError_Success;
    }
    /* large states are not gathered on one rank: they are partitioned over
     * the ranks (every rank then flushes its own, disjoint share) */
    (void)root;
    return pgs_preagg_merge_exchange(s, nccl_comm, rank, nranks);
}

/* ------------------------------------------------------------------
 * Merge of small states over NVLink peer memory - no collective library in
 * the data path (kern_gpupreagg.cuh: gpupreagg_peer_push / _pull).  Every
 * rank allocates an exchange area; only the root's is used.  One process per
 * GPU: the root's area travels as a CUDA IPC handle (the launcher carries its
 * 64 bytes to the ranks, like the NCCL id); several sessions of one process
 * attach by pointer.
 * ------------------------------------------------------------------ */
#define PGS_EXCHANGE_MAX_RANKS  16
#define PGS_PEER_HEAD_WORDS     (16 + 16 * PGS_EXCHANGE_MAX_RANKS)

extern "C" int
pgs_preagg_peer_setup(pgs_session *s, int rank, int nranks, int root, void *ipc_handle_64)
{
    if (refuse_keyheap(s, "pgs_preagg_peer_setup") != StromError_Success)
        return StromError_BadRequestMessage;
    if (!s || nranks < 1 || nranks > PGS_EXCHANGE_MAX_RANKS || rank < 0 || rank >= nranks ||
        root < 0 || root >= nranks)
    {
        set_error("pgs_preagg_peer_setup: bad arguments (at most %d ranks)", PGS_EXCHANGE_MAX_RANKS);
        return StromError_BadRequestMessage;
    }
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    if (s->peer_area)
    {
        set_error("pgs_preagg_peer_setup: the session already has an exchange area");
        return StromError_BadRequestMessage;
    }
    size_t W = s->desc.slot_bytes / 8;
    size_t cap = (s->desc.num_keys == 0 ? 1 : std::min<size_t>(s->gs.gh_nslots, 65536));
    size_t words = PGS_PEER_HEAD_WORDS + 2 * (size_t)nranks * (1 + cap) * W;
    CUDA_CHECK(cudaMalloc((void **)&s->peer_area, words * 8));
    CUDA_CHECK(cudaMemset(s->peer_area, 0, words * 8));
    s->peer_area_bytes = words * 8;
    s->peer_rank = rank;
    s->peer_nranks = nranks;
    s->peer_root = root;
    s->peer_cap = (cl_uint)cap;
    s->peer_epoch = 0;
    s->peer_root_area = (rank == root ? s->peer_area : NULL);
    s->peer_mapped = false;
    if (ipc_handle_64)
    {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
        CUDA_CHECK(cudaIpcGetMemHandle(&h, s->peer_area));
        memcpy(ipc_handle_64, &h, 64);
    }
    for (int i = 0; i < 4; i++)
        if (!s->ev_m[i])
            CUDA_CHECK(cudaEventCreate(&s->ev_m[i]));
    return StromError_Success;
}

extern "C" int
pgs_preagg_peer_attach(pgs_session *s, const void *root_ipc_handle_64)
{
    if (!s || !s->peer_area || !root_ipc_handle_64)
    {
        set_error("pgs_preagg_peer_attach: pgs_preagg_peer_setup comes first");
        return StromError_BadRequestMessage;
    }
    if (s->peer_rank == s->peer_root)
        return StromError_Success;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    cudaIpcMemHandle_t h;
    void *p = NULL;
    memcpy(&h, root_ipc_handle_64, 64);
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_root_area = (cl_ulong *)p;
    s->peer_mapped = true;
    return StromError_Success;
}

extern "C" int
pgs_preagg_peer_attach_session(pgs_session *s, pgs_session *root_session)
{
    if (!s || !root_session || !s->peer_area || !root_session->peer_area ||
        root_session->peer_rank != root_session->peer_root ||
        s->peer_nranks != root_session->peer_nranks || s->peer_cap != root_session->peer_cap)
    {
        set_error("pgs_preagg_peer_attach_session: the two sessions were not set up as ranks "
                  "of one merge");
        return StromError_BadRequestMessage;
    }
    if (s == root_session)
        return StromError_Success;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    if (s->ordinal != root_session->ordinal)
    {
        cudaError_t e = cudaDeviceEnablePeerAccess(root_session->ordinal, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        {
            set_error("cudaDeviceEnablePeerAccess(%d): %s", root_session->ordinal,
                      cudaGetErrorString(e));
            return StromError_CudaInternal;
        }
        cudaGetLastError();
    }
    s->peer_root_area = root_session->peer_area;
    s->peer_mapped = false;
    return StromError_Success;
}

static void
merge_trace_begin(pgs_session *s)
{
    if (s->perfmon && s->ev_m[0])
    {
        /* the events of the merge before this one have long completed */
        float ms = 0;
        if (s->merge_count > 0 && cudaEventElapsedTime(&ms, s->ev_m[0], s->ev_m[1]) == cudaSuccess)
            s->merge_ms[0] += ms;
        else
            cudaGetLastError();
        cudaEventRecord(s->ev_m[0], s->s_exec);
    }
}

static void
merge_trace_end(pgs_session *s)
{
    if (s->perfmon && s->ev_m[1])
        cudaEventRecord(s->ev_m[1], s->s_exec);
    s->merge_count++;
}

/* one merge: every rank calls it once per scan, after its last chunk and
 * before pgs_preagg_finish().  Stream ordered, no host synchronisation, no
 * rendezvous: a rank that is not the root pushes and goes on; the root waits
 * (on the device) for the data only. */
extern "C" int
pgs_preagg_merge_peer(pgs_session *s)
{
    if (refuse_keyheap(s, "pgs_preagg_merge_peer") != StromError_Success)
        return StromError_BadRequestMessage;
    if (!s || !s->peer_area || !s->peer_root_area)
    {
        set_error("pgs_preagg_merge_peer: no exchange area (pgs_preagg_peer_setup / _attach)");
        return StromError_BadRequestMessage;
    }
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    cl_ulong epoch = ++s->peer_epoch;
    cl_uint rank = (cl_uint)s->peer_rank, nranks = (cl_uint)s->peer_nranks;
    cl_uint root = (cl_uint)s->peer_root, cap = s->peer_cap;
    cl_uint *local = (cl_uint *)((char *)s->d_scratch + 80);
    int rc;

    if (nranks == 1)
        return StromError_Success;
    merge_trace_begin(s);
    if (s->peer_rank != s->peer_root)
    {
        if (!s->h_moved)
        {
            CUDA_CHECK(cudaHostAlloc((void **)&s->h_moved, 64,
                                     cudaHostAllocPortable | cudaHostAllocMapped));
            *s->h_moved = 0;
        }
        cl_uint *d_moved = NULL;
        CUDA_CHECK(cudaHostGetDevicePointer((void **)&d_moved, s->h_moved, 0));
        void *args[] = { &s->gs, &s->peer_root_area, &rank, &nranks, &cap, &epoch, &local,
                         &d_moved };
        s->push_pending = true;
        int grid = (s->desc.num_keys == 0) ? 1 :
            std::max(1, std::min<int>(s->num_sms, (int)(((size_t)s->gs.gh_nslots + 255) / 256)));
        rc = launch_kernel(s, s->k_peer_push, grid, 256, 0, args);
    }
    else
    {
        if (!s->d_kg_misc)
            CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
        void *args[] = { &s->gs, &s->peer_root_area, &root, &nranks, &cap, &epoch, &local,
                         &s->d_kg_misc };
        int grid = (s->desc.num_keys == 0) ? 1 :
            std::max(1, std::min<int>(s->num_sms, (int)(((size_t)cap + 255) / 256)));
        if (!s->merge_pending)
            CUDA_CHECK(cudaMemsetAsync(s->d_kg_misc, 0, sizeof(kern_gpupreagg), s->s_exec));
        s->merge_pending = true;
        rc = launch_kernel(s, s->k_peer_pull, grid, 256, 0, args);
    }
    merge_trace_end(s);
    return rc;
}

/* ------------------------------------------------------------------
 * Partitioned exchange of large GROUP BY states: rank r ends
 * up with the groups whose key hash maps to r, every rank then flushes its
 * own share - the ranks' partial rows are disjoint in their keys, and nobody
 * holds (or copies to the host) more than 1/G of the groups.
 *   1. count the records per destination        (gpupreagg_export_parts, pass 0)
 *   2. all ranks learn all counts               (one ncclAllGather, G x G words)
 *   3. write the records bucket by bucket       (gpupreagg_export_parts, pass 1)
 *   4. all-to-all of the buckets                (ncclSend / ncclRecv in one group)
 *   5. reset the state, size the table for what arrives, import
 * One host synchronisation (the counts).  The exchange buffers belong to the
 * session and only ever grow.  Every rank issues every send and receive with
 * the counts all ranks agreed on in step 2, whatever happens locally.
 * ------------------------------------------------------------------ */
static int
import_async(pgs_session *s, const void *device_buf, size_t nrecords)
{
    const cl_ulong *recs = (const cl_ulong *)device_buf;
    while (nrecords > 0)
    {
        cl_uint n = (cl_uint)std::min<size_t>(nrecords, 1u << 30);
        void *args[] = { &s->gs, &recs, &n, &s->d_kg_misc };
        int grid = std::max(1, std::min<int>(s->num_sms * 8, (int)((n + 255) / 256)));
        int rc = launch_kernel(s, s->k_import, grid, 256, 0, args);
        if (rc != StromError_Success)
            return rc;
        recs += (size_t)n * (s->desc.slot_bytes / 8);
        nrecords -= n;
    }
    return StromError_Success;
}

extern "C" int
pgs_preagg_merge_exchange(pgs_session *s, void *nccl_comm, int rank, int nranks)
{
    if (refuse_keyheap(s, "pgs_preagg_merge_exchange") != StromError_Success)
        return StromError_BadRequestMessage;
    const int NCCL_UINT8 = 1;
    if (!s || !nccl_comm || nranks < 1 || nranks > PGS_EXCHANGE_MAX_RANKS ||
        rank < 0 || rank >= nranks)
    {
        set_error("pgs_preagg_merge_exchange: bad arguments (at most %d ranks)",
                  PGS_EXCHANGE_MAX_RANKS);
        return StromError_BadRequestMessage;
    }
    if (s->desc.num_keys == 0)
    {
        set_error("pgs_preagg_merge_exchange: a state without GROUP BY has nothing to partition");
        return StromError_BadRequestMessage;
    }
    int rc = nccl_load();
    if (rc != StromError_Success)
        return rc;
    CUDA_CHECK(cudaSetDevice(s->ordinal));
    if (nranks == 1)
        return StromError_Success;
    rc = drain(s);
    if (rc != StromError_Success)
        return rc;
    const size_t recb = s->desc.slot_bytes;
    const cl_uint R = (cl_uint)nranks;
    cl_uint *d_counts = (cl_uint *)((char *)s->d_scratch + 512);
    cl_uint *d_cursors = d_counts + PGS_EXCHANGE_MAX_RANKS;
    cl_uint *d_offsets = d_cursors + PGS_EXCHANGE_MAX_RANKS;
    cl_uint *d_all = d_offsets + PGS_EXCHANGE_MAX_RANKS;    /* [R][R] */
    cl_ulong *d_send = (cl_ulong *)s->d_xchg;
    int grid = std::max(1, std::min<int>(s->num_sms * 8,
                    (int)std::min<size_t>((state_nslots(s) + s->gs.ovf_cap + 255) / 256, 1u << 30)));
    auto t0 = std::chrono::steady_clock::now();

    if (!s->d_kg_misc)
        CUDA_CHECK(cudaMalloc((void **)&s->d_kg_misc, sizeof(kern_gpupreagg)));
    CUDA_CHECK(cudaMemsetAsync(s->d_kg_misc, 0, sizeof(kern_gpupreagg), s->s_exec));
    s->merge_pending = true;
    CUDA_CHECK(cudaMemsetAsync(d_counts, 0, sizeof(cl_uint) * 3 * PGS_EXCHANGE_MAX_RANKS, s->s_exec));
    {
        cl_uint pass = 0;
        void *args[] = { &s->gs, &d_send, &d_counts, &d_offsets, &d_cursors, (void *)&R, &pass };
        rc = launch_kernel(s, s->k_export_parts, grid, 256, 0, args);
        if (rc != StromError_Success)
            return rc;
    }
    NCCL_CHECK(nccl.allgather(d_counts, d_all, sizeof(cl_uint) * R, NCCL_UINT8, nccl_comm, s->s_exec));
    std::vector<cl_uint> all((size_t)R * R, 0);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, sizeof(cl_uint) * R * R,
                               cudaMemcpyDeviceToHost, s->s_exec));
    CUDA_CHECK(cudaStreamSynchronize(s->s_exec));
    session_bury(s);

    /* Every rank issues every send and receive with the counts agreed on
     * above.  KNOWN GAP: a rank whose local work below fails (device memory
     * for the buffers or the larger table) returns before the exchange and
     * its peers then wait in ncclRecv until the caller aborts the
     * communicator; a second all-gather of the local status in front of the
     * group would close it. */
    std::vector<cl_uint> soff(R + 1, 0), roff(R + 1, 0);
    size_t total_send = 0, total_recv = 0;
    for (cl_uint d = 0; d < R; d++)
    {
        soff[d] = (cl_uint)total_send;
        total_send += all[(size_t)rank * R + d];
    }
    for (cl_uint r = 0; r < R; r++)
    {
        roff[r] = (cl_uint)total_recv;
        if ((int)r != rank)
            total_recv += all[(size_t)r * R + rank];
    }
    if (s->d_xchg_cap < std::max<size_t>(total_send, 1) * recb)
    {
        if (s->d_xchg)
            CUDA_CHECK(cudaFree(s->d_xchg));
        s->d_xchg = NULL;
        s->d_xchg_cap = 0;
        size_t want = std::max<size_t>(total_send, 1) * recb;
        want += want / 8;
        CUDA_CHECK(cudaMalloc((void **)&s->d_xchg, want));
        s->d_xchg_cap = want;
    }
    if (s->d_xrecv_cap < std::max<size_t>(total_recv, 1) * recb)
    {
        if (s->d_xrecv)
            CUDA_CHECK(cudaFree(s->d_xrecv));
        s->d_xrecv = NULL;
        s->d_xrecv_cap = 0;
        size_t want = std::max<size_t>(total_recv, 1) * recb;
        want += want / 8;
        CUDA_CHECK(cudaMalloc((void **)&s->d_xrecv, want));
        s->d_xrecv_cap = want;
    }
    d_send = (cl_ulong *)s->d_xchg;
    CUDA_CHECK(cudaMemcpyAsync(d_offsets, soff.data(), sizeof(cl_uint) * R,
                               cudaMemcpyHostToDevice, s->s_exec));
    {
        cl_uint pass = 1;
        void *args[] = { &s->gs, &d_send, &d_counts, &d_offsets, &d_cursors, (void *)&R, &pass };
        rc = launch_kernel(s, s->k_export_parts, grid, 256, 0, args);
        if (rc != StromError_Success)
            return rc;
    }
    /* the state now lives in the send buffer: start over with an empty one,
     * with a table that takes what this rank will own (every record that
     * arrives may be a group of its own) */
    rc = session_init_state(s);
    if (rc != StromError_Success)
        return rc;
    size_t own = all[(size_t)rank * R + rank];
    rc = session_grow_table(s, 2 * (total_recv + own));
    if (rc != StromError_Success)
        return rc;
    NCCL_CHECK(nccl.group_start());
    for (cl_uint r = 0; r < R; r++)
    {
        if ((int)r == rank)
            continue;
        int e1 = nccl.send((const char *)s->d_xchg + (size_t)soff[r] * recb,
                           (size_t)all[(size_t)rank * R + r] * recb, NCCL_UINT8, (int)r,
                           nccl_comm, s->s_exec);
        int e2 = nccl.recv(s->d_xrecv + (size_t)roff[r] * recb,
                           (size_t)all[(size_t)r * R + rank] * recb, NCCL_UINT8, (int)r,
                           nccl_comm, s->s_exec);
        if (e1 != 0 || e2 != 0)
        {
            nccl.group_end();
            set_error("ncclSend / ncclRecv failed: %s", nccl.errstr ? nccl.errstr(e1 ? e1 : e2) : "?");
            return StromError_CudaInternal;
        }
    }
    NCCL_CHECK(nccl.group_end());
    rc = import_async(s, (const char *)s->d_xchg + (size_t)soff[rank] * recb, own);
    if (rc == StromError_Success)
        rc = import_async(s, s->d_xrecv, total_recv);
    s->merge_count++;
    s->merge_ms[1] += std::chrono::duration<double, std::milli>(
        std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" const char *
pgs_preagg_perfmon_json(pgs_session *s)
{
    pgs::JsonPtr o = pgs::Json::object();
    cl_ulong counters[2] = {0, 0};
    if (s->d_scratch && cudaSetDevice(s->ordinal) == cudaSuccess)
        cudaMemcpy(counters, (char *)s->d_scratch + 16, sizeof(counters), cudaMemcpyDeviceToHost);
    o->set("num_dma_send", (long long)s->num_dma_send);
    o->set("bytes_dma_send", (long long)s->bytes_dma_send);
    o->set("num_dma_recv", (long long)s->num_dma_recv);
    o->set("bytes_dma_recv", (long long)s->bytes_dma_recv);
    o->set("num_chunks", (long long)s->num_chunks);
    o->set("num_rechecked_chunks", (long long)s->num_rechecked_chunks);
    o->set("num_kernel_launches", (long long)s->launches);
    o->set("time_kern_build_ms", pgs::Json::number(s->time_kern_build_ms));
    o->set("time_kern_main_ms", pgs::Json::number(s->time_kern_main_ms));
    o->set("num_kern_main", (long long)s->num_kern_main);
    o->set("rows_kern_main", (long long)s->rows_kern_main);
    o->set("nrows_filtered", (long long)counters[1]);
    if (getenv("PGSTROM_DEBUG_LEVEL"))
    {
        /* cycle counters of GPUPREAGG_DEBUG_LEVEL 4 (kern_gpupreagg.cuh) */
        cl_ulong dbg[4] = {0, 0, 0, 0};
        if (s->d_scratch)
            cudaMemcpy(dbg, (char *)s->d_scratch + 16 + 14 * 8, sizeof(dbg), cudaMemcpyDeviceToHost);
        pgs::JsonPtr a = pgs::Json::array();
        for (int i = 0; i < 4; i++)
            a->push(pgs::Json::number((double)dbg[i]));
        o->set("debug_counters", a);
    }
    o->set("num_table_grown", (long long)s->num_table_grown);
    o->set("merge_count", (long long)s->merge_count);
    o->set("merge_kernel_ms", pgs::Json::number(s->merge_ms[0]));
    o->set("merge_exchange_ms", pgs::Json::number(s->merge_ms[1]));
    o->set("heap_pages_per_stage", (long long)s->heap_pps);
    o->set("heap_num_stages", (long long)s->heap_nstages);
    o->set("grid_heap", s->grid_heap);
    o->set("grid_main", s->grid_main);
    o->set("smem_main", (long long)s->smem_main);
    o->set("sh_nslots", (long long)s->sh_nslots);
    o->set("gh_nslots", (long long)s->gs.gh_nslots);
    o->set("part_nparts", (long long)s->gs.part_nparts);
    o->set("part_cap", (long long)s->gs.part_cap);
    o->set("part_seg_cap", (long long)s->gs.part_seg_cap);
    o->set("part_seg_max", (long long)s->gs.part_seg_max);
    o->set("part_slots", (long long)s->gs.part_slots);
    o->set("block_threads", (long long)s->desc.block_threads);
    o->set("tile_rows", (long long)s->tile_rows);
    o->set("num_stages", (long long)s->nstages);
    o->set("stage_bytes", (long long)(s->tile_rows / 1024) * s->desc.stage_bytes);
    o->set("row_bytes", (long long)s->desc.row_bytes);
    o->set("slot_bytes", (long long)s->desc.slot_bytes);
    o->set("key_heap_nslots", (long long)s->kh.nslots);
    o->set("key_heap_bytes", (long long)s->kh.heap_bytes);
    o->set("key_heap_used", (long long)s->kh_host.size());
    s->perfmon_buf = o->dump();
    return s->perfmon_buf.c_str();
}

extern "C" void
pgs_preagg_abort(pgs_session *s)
{
    if (!s)
        return;
    /* in-flight chunks may still be read by the device: wait for them
     * before the caller frees anything (restrack.c contract) */
    if (cudaSetDevice(s->ordinal) == cudaSuccess)
    {
        if (s->s_copy) cudaStreamSynchronize(s->s_copy);
        if (s->s_exec) cudaStreamSynchronize(s->s_exec);
    }
    for (auto &sl : s->slots)
        sl.busy = false;
    s->aborted = true;
}

extern "C" void
pgs_preagg_close(pgs_session *s)
{
    if (!s)
        return;
    if (cudaSetDevice(s->ordinal) == cudaSuccess)
    {
        if (s->s_copy) cudaStreamSynchronize(s->s_copy);
        if (s->s_exec) cudaStreamSynchronize(s->s_exec);
        for (auto &sl : s->slots)
        {
            if (sl.d_kds) cudaFree(sl.d_kds);
            if (sl.d_kgpreagg) cudaFree(sl.d_kgpreagg);
            if (sl.h_kgpreagg) cudaFreeHost(sl.h_kgpreagg);
            if (sl.d_recheck) cudaFree(sl.d_recheck);
            if (sl.d_heap_index) cudaFree(sl.d_heap_index);
            if (sl.h_status) cudaFreeHost(sl.h_status);
            if (sl.ev_copied) cudaEventDestroy(sl.ev_copied);
            if (sl.ev_done) cudaEventDestroy(sl.ev_done);
            if (sl.ev_k0) cudaEventDestroy(sl.ev_k0);
            if (sl.ev_k1) cudaEventDestroy(sl.ev_k1);
        }
        session_bury(s);
        if (s->gs.gh_slots) cudaFree(s->gs.gh_slots);
        if (s->gs.ovf_recs) cudaFree(s->gs.ovf_recs);
        if (s->peer_mapped && s->peer_root_area) cudaIpcCloseMemHandle(s->peer_root_area);
        if (s->peer_area) cudaFree(s->peer_area);
        if (s->d_xrecv) cudaFree(s->d_xrecv);
        for (int i = 0; i < 4; i++)
            if (s->ev_m[i]) cudaEventDestroy(s->ev_m[i]);
        if (s->gs.part_cursor) cudaFree(s->gs.part_cursor);
        if (s->gs.part_nused) cudaFree(s->gs.part_nused);
        if (s->gs.part_recs) cudaFree(s->gs.part_recs);
        if (s->gs.part_images) cudaFree(s->gs.part_images);
        if (s->gs.part_seg_counts) cudaFree(s->gs.part_seg_counts);
        if (s->kh.slots) cudaFree(s->kh.slots);
        if (s->d_result) cudaFree(s->d_result);
        if (s->d_kg_misc) cudaFree(s->d_kg_misc);
        if (s->h_result_head) cudaFreeHost(s->h_result_head);
        if (s->h_moved) cudaFreeHost(s->h_moved);
        if (s->d_xchg) cudaFree(s->d_xchg);
        if (s->d_scratch) cudaFree(s->d_scratch);
        if (s->s_copy) cudaStreamDestroy(s->s_copy);
        if (s->s_exec) cudaStreamDestroy(s->s_exec);
        if (s->library) cudaLibraryUnload(s->library);
    }
    delete s;
}

extern "C" int
pgs_preagg_key_heap(pgs_session *s, const void **heap, size_t *heap_len)
{
    if (!s || !heap || !heap_len)
    {
        set_error("pgs_preagg_key_heap: bad arguments");
        return StromError_BadRequestMessage;
    }
    *heap = s->kh_host.empty() ? NULL : s->kh_host.data();
    *heap_len = s->kh_host.size();
    return StromError_Success;
}

extern "C" void *
pgs_preagg_stream(pgs_session *s)
{
    return s ? (void *)s->s_exec : NULL;
}

extern "C" uint64_t
pgs_preagg_launch_count(pgs_session *s)
{
    return s ? s->launches : 0;
}

extern "C" int
pgs_device_l2_flush(int device)
{
    /* write a buffer larger than L2 (126 MB on B200) */
    static std::map<int, void *> bufs;
    int ord = device_ordinal(device);
    const size_t sz = 256UL << 20;
    if (ord < 0)
        return StromError_ServerNotReady;
    CUDA_CHECK(cudaSetDevice(ord));
    if (!bufs.count(ord))
    {
        void *p = NULL;
        CUDA_CHECK(cudaMalloc(&p, sz));
        bufs[ord] = p;
    }
    static int v = 0;
    CUDA_CHECK(cudaMemset(bufs[ord], ++v & 0xff, sz));
    CUDA_CHECK(cudaDeviceSynchronize());
    return StromError_Success;
}
