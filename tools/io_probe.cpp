// io_probe: how fast can THIS box take a large text file?  Decides the snps.csv writer's strategy.
//   io_probe <dir> <GB> <threads>
// Writes <GB> GB of prepared text into <dir>/io_probe.tmp five ways and prints GB/s for each:
//   write1    one thread, write() of 64 MB chunks (the round-1 writer)
//   pwriteK   K threads, pwrite() of disjoint 16 MB chunks into a pre-sized file
//   mmapK     K threads memcpy into a shared mapping of the pre-sized file
//   mmapF/mmapP  the same after fallocate / with MADV_POPULATE_WRITE per chunk
//   directK   K threads, O_DIRECT pwrite() of 16 MB aligned chunks into an fallocate'd file
//   filesK    K threads, one file each (upper bound without the shared inode)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statfs.h>
#include <unistd.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: io_probe <dir> <GB> <threads>\n"); return 2; }
    const std::string dir = argv[1];
    const size_t total = size_t(atof(argv[2]) * 1e9) / (16 << 20) * (16 << 20);
    const int K = atoi(argv[3]);
    const size_t chunk = 16 << 20;
    char *src = nullptr;
    if (posix_memalign(reinterpret_cast<void **>(&src), 4096, 256 << 20) != 0) return 1;
    for (size_t i = 0; i < (256u << 20); ++i) src[i] = "0123456789,\n"[i % 12];
    const std::string path = dir + "/io_probe.tmp";
    struct statfs sf;
    if (statfs(dir.c_str(), &sf) == 0) printf("dir %s  f_type 0x%lx  free %.1f GB\n", dir.c_str(), (unsigned long)sf.f_type, double(sf.f_bavail) * sf.f_bsize / 1e9);
    auto report = [&](const char *what, double dt) { printf("%-10s %6.2f s  %6.2f GB/s\n", what, dt, total / dt / 1e9); fflush(stdout); };
    auto run_threads = [&](auto fn) { std::vector<std::thread> p; for (int t = 0; t < K; ++t) p.emplace_back(fn, t); for (auto &th : p) th.join(); };
    const size_t chunks = total / chunk;
    {   // write1
        unlink(path.c_str());
        int fd = open(path.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644);
        double t0 = now();
        for (size_t c = 0; c < chunks; ++c) { const char *p = src + (c % 16) * chunk; size_t left = chunk; while (left) { ssize_t k = write(fd, p, left); if (k <= 0) { perror("write"); return 1; } p += k; left -= k; } }
        close(fd);
        report("write1", now() - t0);
    }
    {   // pwriteK
        unlink(path.c_str());
        int fd = open(path.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644);
        double t0 = now();
        if (ftruncate(fd, off_t(total)) != 0) perror("ftruncate");
        run_threads([&](int t) { for (size_t c = t; c < chunks; c += K) { const char *p = src + (c % 16) * chunk; size_t left = chunk; off_t off = off_t(c * chunk); while (left) { ssize_t k = pwrite(fd, p, left, off); if (k <= 0) { perror("pwrite"); return; } p += k; left -= k; off += k; } } });
        close(fd);
        report("pwriteK", now() - t0);
    }
    {   // mmapK
        unlink(path.c_str());
        int fd = open(path.c_str(), O_CREAT | O_RDWR | O_TRUNC, 0644);
        double t0 = now();
        if (ftruncate(fd, off_t(total)) != 0) perror("ftruncate");
        char *m = static_cast<char *>(mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
        if (m == MAP_FAILED) { perror("mmap"); } else {
            run_threads([&](int t) { for (size_t c = t; c < chunks; c += K) memcpy(m + c * chunk, src + (c % 16) * chunk, chunk); });
            munmap(m, total);
        }
        close(fd);
        report("mmapK", now() - t0);
    }
    for (int variant = 1; variant <= 2; ++variant) {   // mmapF: blocks allocated first; mmapP: + MADV_POPULATE_WRITE per chunk
        unlink(path.c_str());
        int fd = open(path.c_str(), O_CREAT | O_RDWR | O_TRUNC, 0644);
        double t0 = now();
        if (fallocate(fd, 0, 0, off_t(total)) != 0) { if (ftruncate(fd, off_t(total)) != 0) perror("ftruncate"); }
        char *m = static_cast<char *>(mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
        if (m == MAP_FAILED) { perror("mmap"); } else {
            run_threads([&](int t) { for (size_t c = t; c < chunks; c += K) {
#ifdef MADV_POPULATE_WRITE
                if (variant == 2) madvise(m + c * chunk, chunk, MADV_POPULATE_WRITE);
#endif
                memcpy(m + c * chunk, src + (c % 16) * chunk, chunk); } });
            munmap(m, total);
        }
        close(fd);
        report(variant == 1 ? "mmapF" : "mmapP", now() - t0);
    }
    {   // directK
        unlink(path.c_str());
        int fd = open(path.c_str(), O_CREAT | O_WRONLY | O_TRUNC | O_DIRECT, 0644);
        if (fd < 0) { printf("directK    O_DIRECT not supported here (%s)\n", strerror(errno)); }
        else {
            double t0 = now();
            if (fallocate(fd, 0, 0, off_t(total)) != 0) perror("fallocate");
            run_threads([&](int t) { for (size_t c = t; c < chunks; c += K) { const char *p = src + (c % 16) * chunk; size_t left = chunk; off_t off = off_t(c * chunk); while (left) { ssize_t k = pwrite(fd, p, left, off); if (k <= 0) { perror("pwrite O_DIRECT"); return; } p += k; left -= k; off += k; } } });
            close(fd);
            report("directK", now() - t0);
        }
    }
    {   // filesK
        double t0 = now();
        run_threads([&](int t) { std::string p = path + "." + std::to_string(t); int fd = open(p.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644); for (size_t c = t; c < chunks; c += K) { const char *q = src + (c % 16) * chunk; size_t left = chunk; while (left) { ssize_t k = write(fd, q, left); if (k <= 0) return; q += k; left -= k; } } close(fd); });
        report("filesK", now() - t0);
        for (int t = 0; t < K; ++t) unlink((path + "." + std::to_string(t)).c_str());
    }
    unlink(path.c_str());
    return 0;
}
