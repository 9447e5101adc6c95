// multi.cpp — the multi-device entry points of include/ftsb.h, written on top of the single-device C-ABI.
//
// Instances (and .dc chains) are independent and share one compiled topology, so a batch is cut into
// contiguous shards, one per device of the list; every device has its own handle and its own persistent host thread
// (pinned to the cores of the GPU's NUMA node), which calls the single-device entry points on its shard with pointers
// into the CALLER's buffers: results land where a single-device call would have put them, there is no gather step and
// no collective.  The caller — the reference's `main` (src/main.rs:29-44) through the Rust binding — needs no torch, no
// process group and no knowledge of the device count.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ftsb.h"

namespace {

// cores of the NUMA node the device hangs on ("" when unknown)
std::vector<int> numa_cpus_of_device(int device)
{
    std::vector<int> cpus;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return cpus;
    }
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[256];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node < 0) return cpus;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    char list[4096] = {0};
    if (FILE *f = fopen(path, "r")) {
        if (!fgets(list, sizeof list, f)) list[0] = 0;
        fclose(f);
    }
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) {
            for (int i = a; i <= b; ++i) cpus.push_back(i);
        } else if (sscanf(tok, "%d", &a) == 1)
            cpus.push_back(a);
    }
    return cpus;
}

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = false, quit = false;
    int rc = 0;
};

} // namespace

struct ftsb_multi {
    std::vector<ftsb_circuit *> h;
    std::vector<int> dev;
    std::vector<Worker *> w;
    std::string err, variant;
    int numa_pin = 1;
};

namespace {

thread_local std::string g_multi_error;

void worker_main(ftsb_multi *m, int g)
{
    Worker &w = *m->w[g];
    cudaSetDevice(m->dev[g]);
    if (m->numa_pin) {
        const std::vector<int> cpus = numa_cpus_of_device(m->dev[g]);
        if (!cpus.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus)
                if (c < CPU_SETSIZE) CPU_SET(c, &set);
            pthread_setaffinity_np(pthread_self(), sizeof set, &set); // best effort
        }
    }
    for (;;) {
        std::function<int()> t;
        {
            std::unique_lock<std::mutex> lk(w.mu);
            w.cv.wait(lk, [&] { return w.has_task || w.quit; });
            if (w.quit) return;
            t = w.task;
            w.has_task = false;
        }
        const int rc = t();
        {
            std::lock_guard<std::mutex> lk(w.mu);
            w.rc = rc;
            w.done = true;
        }
        w.cv.notify_all();
    }
}

// runs fn(g, first, count) on every device's worker for its shard of [0, batch); first error wins
int run_sharded(ftsb_multi *m, int64_t batch, const std::function<int(int, int64_t, int64_t)> &fn)
{
    const int G = (int)m->h.size();
    const int64_t per = (batch + G - 1) / G;
    std::vector<char> used(G, 0);
    for (int g = 0; g < G; ++g) {
        const int64_t first = std::min<int64_t>((int64_t)g * per, batch), count = std::max<int64_t>(0, std::min(per, batch - first));
        if (count == 0) continue;
        Worker &w = *m->w[g];
        {
            std::lock_guard<std::mutex> lk(w.mu);
            w.task = [=] { return fn(g, first, count); };
            w.has_task = true;
            w.done = false;
        }
        w.cv.notify_all();
        used[g] = 1;
    }
    int rc = 0;
    for (int g = 0; g < G; ++g) {
        if (!used[g]) continue;
        Worker &w = *m->w[g];
        std::unique_lock<std::mutex> lk(w.mu);
        w.cv.wait(lk, [&] { return w.done; });
        if (w.rc && !rc) {
            rc = w.rc;
            m->err = "device " + std::to_string(m->dev[g]) + ": " + ftsb_last_error(m->h[g]);
        }
    }
    if (!rc) {
        m->variant = std::to_string(G) + " devices x [" + ftsb_last_variant(m->h[0]) + "]";
    }
    return rc;
}

int bad(ftsb_multi *m, const char *msg)
{
    if (m)
        m->err = msg;
    else
        g_multi_error = msg;
    return FTSB_E_ARG;
}

} // namespace

extern "C" {

const char *ftsb_multi_last_error(const ftsb_multi *m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

int ftsb_multi_compile(const ftsb_elem *elems, int32_t n_elems, int32_t n_run, int32_t n_start, const int32_t *row_class,
                       const int32_t *devices, int32_t n_devices, ftsb_multi **out)
{
    if (!out) return bad(nullptr, "ftsb_multi_compile: out is NULL");
    *out = nullptr;
    std::vector<int> devs;
    if (devices && n_devices > 0)
        devs.assign(devices, devices + n_devices);
    else { // every visible device
        const int n = ftsb_device_count();
        for (int i = 0; i < n; ++i) devs.push_back(i);
    }
    if (devs.empty()) {
        g_multi_error = "ftsb_multi_compile: no usable CUDA device; this library has no CPU fallback";
        return FTSB_E_CUDA;
    }
    ftsb_multi *m = new ftsb_multi();
    if (const char *e = getenv("FTSB_NUMA_PIN")) m->numa_pin = atoi(e) ? 1 : 0; // the workers read it when they start
    for (int d : devs) {
        ftsb_circuit *c = nullptr;
        const int rc = ftsb_compile(elems, n_elems, n_run, n_start, row_class, d, &c);
        if (rc) {
            g_multi_error = std::string("device ") + std::to_string(d) + ": " + ftsb_last_error(nullptr);
            for (ftsb_circuit *p : m->h) ftsb_destroy(p);
            delete m;
            return rc;
        }
        m->h.push_back(c);
        m->dev.push_back(d);
    }
    for (size_t g = 0; g < m->h.size(); ++g) m->w.push_back(new Worker());
    for (size_t g = 0; g < m->h.size(); ++g) m->w[g]->th = std::thread(worker_main, m, (int)g);
    *out = m;
    return 0;
}

void ftsb_multi_destroy(ftsb_multi *m)
{
    if (!m) return;
    for (Worker *w : m->w) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (ftsb_circuit *c : m->h) ftsb_destroy(c);
    delete m;
}

int32_t ftsb_multi_device_count(const ftsb_multi *m) { return m ? (int32_t)m->h.size() : 0; }
const char *ftsb_multi_last_variant(const ftsb_multi *m) { return m ? m->variant.c_str() : ""; }

int ftsb_multi_set_option(ftsb_multi *m, const char *name, int64_t value)
{
    if (!m || !name) return bad(m, "ftsb_multi_set_option: null argument");
    for (ftsb_circuit *c : m->h) {
        const int rc = ftsb_set_option(c, name, value);
        if (rc) {
            m->err = ftsb_last_error(c);
            return rc;
        }
    }
    return 0;
}

int ftsb_multi_solve_op(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, double *x, int32_t *n_iters,
                        int32_t *status)
{
    if (!m) return bad(nullptr, "null handle");
    if (batch <= 0 || !x || !n_iters || !status) return bad(m, "ftsb_multi_solve_op: bad argument");
    const int ne = ftsb_n_elems(m->h[0]), nf = ftsb_n_fn(m->h[0]), n = ftsb_n_start(m->h[0]);
    return run_sharded(m, batch, [=](int g, int64_t first, int64_t count) {
        return ftsb_solve_op(m->h[g], count, vals ? vals + first * ne : nullptr, fn ? fn + first * nf * 7 : nullptr,
                             FTSB_LAYOUT_INSTANCE_MAJOR, x + first * n, n_iters + first, status + first);
    });
}

int ftsb_multi_solve_dc(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, int32_t sweep_elem, const double *start,
                        const double *stop, const double *step, int64_t max_points, double *x, int32_t *n_iters,
                        int64_t *points_done, int32_t *status)
{
    if (!m) return bad(nullptr, "null handle");
    if (batch <= 0 || !start || !stop || !step || !x || !n_iters || !points_done || !status || max_points <= 0)
        return bad(m, "ftsb_multi_solve_dc: bad argument");
    const int ne = ftsb_n_elems(m->h[0]), nf = ftsb_n_fn(m->h[0]), n = ftsb_n_run(m->h[0]);
    return run_sharded(m, batch, [=](int g, int64_t first, int64_t count) {
        return ftsb_solve_dc(m->h[g], count, vals ? vals + first * ne : nullptr, fn ? fn + first * nf * 7 : nullptr,
                             FTSB_LAYOUT_INSTANCE_MAJOR, sweep_elem, start + first, stop + first, step + first, max_points,
                             x + first * max_points * n, n_iters + first * max_points, points_done + first, status + first);
    });
}

int ftsb_multi_solve_tran(ftsb_multi *m, int64_t batch, const double *vals, const double *fn, double stop, double step_max,
                          const ftsb_tran_opts *opts, double *t, double *x, int32_t *n_iters, int64_t *n_rec, double *x_final,
                          int32_t *status)
{
    if (!m) return bad(nullptr, "null handle");
    if (batch <= 0 || !n_rec || !status) return bad(m, "ftsb_multi_solve_tran: bad argument");
    const int ne = ftsb_n_elems(m->h[0]), nf = ftsb_n_fn(m->h[0]), n = ftsb_n_run(m->h[0]);
    const int64_t cap = opts ? opts->capacity : 0;
    const ftsb_tran_opts o = opts ? *opts : ftsb_tran_opts{0, 1, 0};
    return run_sharded(m, batch, [=](int g, int64_t first, int64_t count) {
        int rc = ftsb_set_params(m->h[g], count, vals ? vals + first * ne : nullptr, fn ? fn + first * nf * 7 : nullptr,
                                 FTSB_LAYOUT_INSTANCE_MAJOR, FTSB_MEM_HOST);
        if (rc) return rc;
        return ftsb_run_tran(m->h[g], stop, step_max, &o, t ? t + first * cap : nullptr, x ? x + first * cap * n : nullptr,
                             n_iters ? n_iters + first * cap : nullptr, n_rec + first, x_final ? x_final + first * n : nullptr,
                             status + first, FTSB_MEM_HOST);
    });
}

int ftsb_multi_get_counters(ftsb_multi *m, ftsb_counters *out)
{
    if (!m || !out) return bad(m, "ftsb_multi_get_counters: null argument");
    memset(out, 0, sizeof *out);
    for (ftsb_circuit *c : m->h) {
        ftsb_counters k;
        const int rc = ftsb_get_counters(c, &k);
        if (rc) {
            m->err = ftsb_last_error(c);
            return rc;
        }
        out->solve_points += k.solve_points;
        out->newton_iters += k.newton_iters;
        out->lu_factor += k.lu_factor;
        out->lu_solve += k.lu_solve;
        out->rej_newton += k.rej_newton;
        out->rej_plte += k.rej_plte;
        out->redo += k.redo;
        out->sparse_flops += k.sparse_flops;
    }
    return 0;
}

int64_t ftsb_multi_launch_count(const ftsb_multi *m)
{
    int64_t n = 0;
    if (m)
        for (ftsb_circuit *c : m->h) n += ftsb_launch_count(c);
    return n;
}

} // extern "C"
