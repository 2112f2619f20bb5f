// Host<->device copy ceiling of a multi-GPU box, per GPU SET and per host-buffer placement: what bounds bench.py's `e2e`
// (frames come from host memory) when several ranks copy at once.  One host thread per GPU of the set, all copying
// `bytes` per pass from their own page-locked buffer at the same time.
//
//   nvcc -O2 -std=c++17 -o aruco3_b200/csrc/build/pcie_matrix tools/pcie_matrix.cu -lpthread
//   pcie_matrix [--mb 1024] [--passes 4] SET [SET ...]      SET = comma-separated device indices, e.g. 0 0,1 0,1,2,3 4,5,6,7
//
// Modes (each set is measured in every mode):
//   plain     cudaHostAlloc(default), allocated and first touched by an unbound thread
//   local     the thread is first bound to the cores the GPU's PCI device lists as local (/sys/bus/pci/devices/<id>/local_cpulist),
//             then allocates and touches: the pages land on the GPU's NUMA node
//   remote    bound to cores NOT in that list (the other socket), to show what a wrong placement costs
//   wc        cudaHostAllocWriteCombined, bound like `local`
// Output: one JSON object per (set, mode, direction) on stdout.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

static std::vector<int> parse_cpulist(const std::string &s) {
    std::vector<int> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        const size_t dash = tok.find('-');
        const int a = atoi(tok.substr(0, dash).c_str()), b = dash == std::string::npos ? a : atoi(tok.substr(dash + 1).c_str());
        for (int i = a; i <= b; i++) v.push_back(i);
    }
    return v;
}

static std::string sysfs(int dev, const char *leaf) {
    char id[64] = {0};
    cudaDeviceGetPCIBusId(id, sizeof(id), dev);
    for (char *p = id; *p; p++) *p = (char)tolower(*p);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + id + "/" + leaf);
    std::string s;
    std::getline(f, s);
    return s;
}

static bool bind(const std::vector<int> &cpus) {
    if (cpus.empty()) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

struct Barrier {
    std::atomic<int> count{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait(int phase) {
        count.fetch_add(1);
        while (count.load() < n * phase) std::this_thread::yield();
    }
};

int main(int argc, char **argv) {
    size_t mb = 1024;
    int passes = 4;
    std::vector<std::vector<int>> sets;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--mb") && i + 1 < argc) mb = (size_t)atol(argv[++i]);
        else if (!strcmp(argv[i], "--passes") && i + 1 < argc) passes = atoi(argv[++i]);
        else sets.push_back(parse_cpulist(argv[i]));
    }
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (sets.empty())
        for (int d = 0; d < ndev; d++) sets.push_back({d});
    const size_t bytes = mb << 20;
    const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t all;
    sched_getaffinity(0, sizeof(all), &all);
    for (int d = 0; d < ndev; d++)
        printf("{\"device\": %d, \"pci\": \"%s\", \"numa_node\": \"%s\", \"local_cpulist\": \"%s\", \"cpus_online\": %ld, \"cpus_allowed\": %d}\n", d,
               sysfs(d, "uevent").c_str(), sysfs(d, "numa_node").c_str(), sysfs(d, "local_cpulist").c_str(), ncpu, CPU_COUNT(&all));
    const char *modes[] = {"plain", "local", "remote", "wc"};
    for (const auto &set : sets) {
        bool ok = true;
        for (int d : set) ok = ok && d >= 0 && d < ndev;
        if (!ok) continue;
        for (int mode = 0; mode < 4; mode++)
            for (int dir = 0; dir < 2; dir++) {  // 0 = host to device, 1 = device to host
                const int n = (int)set.size();
                std::vector<double> gbs(n, 0.0);
                std::vector<int> bound(n, 0);
                Barrier bar(n);
                std::vector<std::thread> th;
                std::atomic<long long> t_first{0}, t_last{0};
                for (int k = 0; k < n; k++)
                    th.emplace_back([&, k] {
                        const int dev = set[k];
                        cudaSetDevice(dev);
                        sched_setaffinity(0, sizeof(all), &all);
                        const std::vector<int> local = parse_cpulist(sysfs(dev, "local_cpulist"));
                        if (mode == 1 || mode == 3) bound[k] = bind(local);
                        if (mode == 2) {
                            std::vector<int> other;
                            for (int c = 0; c < (int)ncpu; c++) {
                                bool in = false;
                                for (int l : local) in = in || l == c;
                                if (!in && CPU_ISSET(c, &all)) other.push_back(c);
                            }
                            bound[k] = bind(other);
                        }
                        void *h = nullptr, *d = nullptr;
                        if (cudaHostAlloc(&h, bytes, mode == 3 ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess ||
                            cudaMalloc(&d, bytes) != cudaSuccess) {
                            fprintf(stderr, "allocation failed on device %d\n", dev);
                            exit(1);
                        }
                        memset(h, k + 1, bytes);  // first touch
                        cudaStream_t s;
                        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                        cudaEvent_t e0, e1;
                        cudaEventCreate(&e0);
                        cudaEventCreate(&e1);
                        auto copy = [&] {
                            if (dir == 0) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
                            else cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
                        };
                        copy();
                        cudaStreamSynchronize(s);
                        bar.wait(1);
                        const long long t0 = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
                        long long expect = 0;
                        t_first.compare_exchange_strong(expect, t0);
                        cudaEventRecord(e0, s);
                        for (int p = 0; p < passes; p++) copy();
                        cudaEventRecord(e1, s);
                        cudaStreamSynchronize(s);
                        const long long t1 = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
                        long long cur = t_last.load();
                        while (t1 > cur && !t_last.compare_exchange_weak(cur, t1)) {}
                        float ms = 0;
                        cudaEventElapsedTime(&ms, e0, e1);
                        gbs[k] = (double)bytes * passes / (ms * 1e-3) / 1e9;
                        bar.wait(2);
                        cudaFreeHost(h);
                        cudaFree(d);
                        cudaStreamDestroy(s);
                        cudaEventDestroy(e0);
                        cudaEventDestroy(e1);
                    });
                for (auto &t : th) t.join();
                double sum = 0, mn = 1e30;
                std::string per = "[", gl = "[";
                for (int k = 0; k < n; k++) {
                    sum += gbs[k];
                    mn = gbs[k] < mn ? gbs[k] : mn;
                    char buf[32];
                    snprintf(buf, sizeof(buf), "%s%.1f", k ? ", " : "", gbs[k]);
                    per += buf;
                    snprintf(buf, sizeof(buf), "%s%d", k ? ", " : "", set[k]);
                    gl += buf;
                }
                const double wall = (double)(t_last.load() - t_first.load()) * 1e-9;
                printf("{\"gpus\": %s], \"mode\": \"%s\", \"dir\": \"%s\", \"mb_per_pass\": %zu, \"passes\": %d, \"per_gpu_gbs\": %s], \"sum_gbs\": %.1f, "
                       "\"min_gbs\": %.1f, \"aggregate_wall_gbs\": %.1f, \"bound\": %d}\n",
                       gl.c_str(), modes[mode], dir ? "d2h" : "h2d", mb, passes, per.c_str(), sum, mn, (double)bytes * passes * n / wall / 1e9, bound[0]);
                fflush(stdout);
            }
    }
    return 0;
}
