/*
 * main.cpp - GUC table, error strings and other process-wide bits.
 *
 * Mirrors main.c of the reference: GUC names, defaults and ranges are kept
 * verbatim (main.c:104-234; gpupreagg.c:2946-2967; gpuscan.c:1704;
 * gpuhashjoin.c:4056; mqueue.c:740; opencl_devprog.c:929-948;
 * shmem.c:1432-1452; opencl_serv.c:408; opencl_devinfo.c:1024-1070) so that
 * an existing postgresql.conf still loads.  GUCs of replaced subsystems
 * (OpenCL server, shared-memory zones, message queue) are accepted and
 * either ignored or re-mapped; `desc` says which.
 */
#include <cstdio>
#include <cstring>
#include <mutex>
#include "pgs_plan.h"
#include "../../include/pgstrom_cuda.h"

namespace pgs {

static std::mutex guc_lock;

static std::vector<GucEntry>
make_guc_table()
{
    std::vector<GucEntry> t = {
        {"pg_strom.enabled", "bool", "on", "on", NULL, NULL, "USERSET",
         "Enables the planner's use of PG-Strom"},
        {"pg_strom.enabled_global", "bool", "on", "on", NULL, NULL, "SUSET",
         "Enables the planner's use of PG-Strom in global"},
        {"pg_strom.perfmon", "bool", "off", "off", NULL, NULL, "USERSET",
         "Enables the performance monitor of PG-Strom"},
        {"pg_strom.show_device_kernel", "bool", "off", "off", NULL, NULL, "USERSET",
         "Enables to show device kernel on EXPLAIN"},
        {"pg_strom.chunk_size", "int", "15", "15", "4", "128", "USERSET",
         "default size of pgstrom_data_store in MB"},
        {"pg_strom.min_async_chunks", "int", "2", "2", "2", "2147483647", "USERSET",
         "least number of chunks to be run asynchronously"},
        {"pg_strom.max_async_chunks", "int", "3", "3", "3", "2147483647", "USERSET",
         "max number of chunk to be run asynchronously"},
        {"gpu_setup_cost", "real", "500", "500", "0", NULL, "USERSET",
         "Cost to setup GPU device to run"},
        {"gpu_operator_cost", "real", "0.000025", "0.000025", "0", NULL, "USERSET",
         "Cost of processing each operators by GPU"},
        {"gpu_tuple_cost", "real", "0.0003125", "0.0003125", "0", NULL, "USERSET",
         "Cost of processing each tuple for GPU"},
        {"enable_gpupreagg", "bool", "on", "on", NULL, NULL, "USERSET",
         "Enables the use of GPU preprocessed aggregate"},
        {"pg_strom.debug_force_gpupreagg", "bool", "off", "off", NULL, NULL, "USERSET",
         "Force GpuPreAgg regardless of the cost (debug)"},
        {"enable_gpuscan", "bool", "on", "on", NULL, NULL, "USERSET",
         "Enables the use of GPU accelerated full-scan"},
        {"enable_gpuhashjoin", "bool", "on", "on", NULL, NULL, "USERSET",
         "Enables the use of GPU accelerated hash-join (accepted; operator not built)"},
        {"pg_strom.key_heap_size", "int", "64", "64", "0", "65536", "USERSET",
         "size of the device heap of long text grouping keys in MB (0: such rows are re-checked on the host)"},
        {"pg_strom.mqueue_timeout", "int", "60000", "60000", "1", "2147483647", "POSTMASTER",
         "timeout of device completion wait in ms (was: message queue wait)"},
        {"pg_strom.devprog_enable_optimization", "bool", "on", "on", NULL, NULL, "SIGHUP",
         "enables optimization on device program build (NVRTC -O3 vs -G-less -O0)"},
        {"pg_strom.devprog_reclaim_threshold", "int", "16384", "16384", "0", "2147483647", "POSTMASTER",
         "threshold of the device program cache to reclaim unused ones, in KB"},
        {"pg_strom.shmem_totalsize", "int", "2048", "2048", "512", "2147483647", "POSTMASTER",
         "total size of pinned host memory for chunks in MB (was: shared memory zones)"},
        {"pg_strom.shmem_maxzones", "int", "256", "256", "1", "1024", "POSTMASTER",
         "accepted and ignored (no zone allocator in the CUDA layer)"},
        {"pg_strom.opencl_num_threads", "int", "0", "0", "0", "2147483647", "POSTMASTER",
         "accepted and ignored (no server threads; one stream set per session)"},
        {"pg_strom.opencl_platform", "int", "-1", "-1", "-1", "2147483647", "POSTMASTER",
         "accepted and ignored (single CUDA platform)"},
        {"pg_strom.opencl_devices", "string", "any", "any", NULL, NULL, "POSTMASTER",
         "CUDA devices to be used: 'any' or a comma separated list of ordinals"},
        {"pg_strom.opencl_device_types", "string", "gpu,accelerator", "gpu,accelerator", NULL, NULL, "POSTMASTER",
         "accepted and ignored (CUDA devices are GPUs)"},
    };
    return t;
}

std::vector<GucEntry> &
guc_table()
{
    static std::vector<GucEntry> table = make_guc_table();
    return table;
}

static bool
parse_bool(const std::string &v, bool *out)
{
    std::string s;
    for (char c : v) s += (char)tolower(c);
    if (s == "on" || s == "true" || s == "t" || s == "yes" || s == "1")
    { *out = true; return true; }
    if (s == "off" || s == "false" || s == "f" || s == "no" || s == "0")
    { *out = false; return true; }
    return false;
}

bool
guc_set(const std::string &name, const std::string &value, std::string *err)
{
    std::lock_guard<std::mutex> g(guc_lock);
    for (GucEntry &e : guc_table())
    {
        if (name != e.name)
            continue;
        std::string kind = e.kind;
        if (kind == "bool")
        {
            bool b;
            if (!parse_bool(value, &b))
            {
                if (err) *err = "parameter \"" + name + "\" requires a Boolean value";
                return false;
            }
            e.value = b ? "on" : "off";
        }
        else if (kind == "int")
        {
            char *end;
            long long v = strtoll(value.c_str(), &end, 10);
            if (*end != '\0' || end == value.c_str())
            {
                if (err) *err = "invalid value for parameter \"" + name + "\": \"" + value + "\"";
                return false;
            }
            if ((e.minval && v < atoll(e.minval)) || (e.maxval && v > atoll(e.maxval)))
            {
                if (err)
                    *err = value + " is outside the valid range for parameter \"" + name +
                        "\" (" + (e.minval ? e.minval : "") + " .. " + (e.maxval ? e.maxval : "") + ")";
                return false;
            }
            e.value = std::to_string(v);
        }
        else if (kind == "real")
        {
            char *end;
            double v = strtod(value.c_str(), &end);
            if (*end != '\0' || end == value.c_str() || (e.minval && v < atof(e.minval)))
            {
                if (err) *err = "invalid value for parameter \"" + name + "\": \"" + value + "\"";
                return false;
            }
            e.value = value;
        }
        else
            e.value = value;
        if (name == "pg_strom.max_async_chunks" || name == "pg_strom.min_async_chunks")
        {
            long long mn = 0, mx = 0;
            for (GucEntry &x : guc_table())
            {
                if (!strcmp(x.name, "pg_strom.min_async_chunks")) mn = atoll(x.value.c_str());
                if (!strcmp(x.name, "pg_strom.max_async_chunks")) mx = atoll(x.value.c_str());
            }
            if (mx <= mn)
            {
                if (err)
                    *err = "\"pg_strom.max_async_chunks\" must be larger than \"pg_strom.min_async_chunks\"";
                /* keep the assignment like PostgreSQL would refuse at startup only */
            }
        }
        return true;
    }
    if (err) *err = "unrecognized configuration parameter \"" + name + "\"";
    return false;
}

std::string
guc_get(const std::string &name)
{
    std::lock_guard<std::mutex> g(guc_lock);
    for (GucEntry &e : guc_table())
        if (name == e.name)
            return e.value;
    return "";
}

bool guc_bool(const std::string &name) { return guc_get(name) == "on"; }
long long guc_int(const std::string &name) { return atoll(guc_get(name).c_str()); }
double guc_real(const std::string &name) { return atof(guc_get(name).c_str()); }

void
guc_reset_all()
{
    std::lock_guard<std::mutex> g(guc_lock);
    for (GucEntry &e : guc_table())
        e.value = e.boot;
}

/* wrapper of pg_strom.enabled and pg_strom.enabled_global (main.c:64-76) */
bool
pgstrom_enabled()
{
    return guc_bool("pg_strom.enabled") && guc_bool("pg_strom.enabled_global");
}

thread_local std::string last_error;

}   /* namespace pgs */

using namespace pgs;

extern "C" {

int
pgstrom_abi_version(void)
{
    return PGSTROM_CUDA_ABI_VERSION;
}

/* translation from StromError_* to human readable form (main.c:288-330) */
const char *
pgstrom_strerror(int errcode)
{
    static thread_local char unknown_buf[256];

    switch (errcode)
    {
        case StromError_Success:            return "Success";
        case StromError_RowFiltered:        return "Row is filtered";
        case StromError_CpuReCheck:         return "To be re-checked by CPU";
        case StromError_ServerNotReady:     return "CUDA device layer is not ready";
        case StromError_BadRequestMessage:  return "Request message is bad";
        case StromError_OpenCLInternal:     return "device runtime internal error";
        case StromError_OutOfSharedMemory:  return "out of shared memory";
        case StromError_OutOfMemory:        return "out of host memory";
        case StromError_DataStoreCorruption:return "data store is corrupted";
        case StromError_DataStoreNoSpace:   return "data store has no space";
        case StromError_DataStoreOutOfRange:return "out of range in data store";
        case StromError_DataStoreReCheck:   return "data store be rechecked";
        case StromError_SanityCheckViolation: return "sanity check violation";
        case StromError_ProgramBuildFailure:return "device program build failure";
        case StromError_CudaInternal:       return "CUDA runtime error";
        default:
            snprintf(unknown_buf, sizeof(unknown_buf),
                     "undefined strom error (code: %d)", errcode);
            break;
    }
    return unknown_buf;
}

const char *
pgs_last_error(void)
{
    return last_error.c_str();
}

int
pgstrom_guc_set(const char *name, const char *value)
{
    std::string err;
    if (!guc_set(name ? name : "", value ? value : "", &err))
    {
        last_error = err;
        return StromError_BadRequestMessage;
    }
    if (!err.empty())
        last_error = err;
    return StromError_Success;
}

const char *
pgstrom_guc_get(const char *name)
{
    static thread_local std::string buf;
    bool found = false;
    for (GucEntry &e : guc_table())
        if (name && !strcmp(name, e.name))
            found = true;
    if (!found)
        return NULL;
    buf = guc_get(name);
    return buf.c_str();
}

const char *
pgstrom_guc_list_json(void)
{
    static thread_local std::string buf;
    JsonPtr arr = Json::array();
    for (GucEntry &e : guc_table())
    {
        JsonPtr o = Json::object();
        o->set("name", e.name);
        o->set("kind", e.kind);
        o->set("value", e.value);
        o->set("boot", e.boot);
        o->set("min", e.minval ? Json::string(e.minval) : Json::null());
        o->set("max", e.maxval ? Json::string(e.maxval) : Json::null());
        o->set("context", e.context);
        o->set("desc", e.desc);
        arr->push(o);
    }
    buf = arr->dump();
    return buf.c_str();
}

void
pgstrom_guc_reset_all(void)
{
    guc_reset_all();
}

}   /* extern "C" */
