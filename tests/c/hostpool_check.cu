// CPU check of the host copy pool (xbitinfo_b200/csrc/xbi_hostpool.cu): contiguous and pitched copies,
// odd sizes and offsets, several threads, and a forked child that starts a pool of its own.
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "xbi_internal.h"

using namespace xbi;

int main() {
    const size_t n = (size_t)96 << 20;
    std::vector<char> a(n), b(n, 0);
    for (size_t i = 0; i < n; ++i) a[i] = (char)((i * 2654435761u) >> 13);
    if (host_copy_threads() < 1) return 1;
    host_copy_parallel(b.data(), n, a.data(), n, n, 1);
    if (memcmp(a.data(), b.data(), n) != 0) return 2;
    // pitched: 5000 rows of 3001 bytes, source pitch 7000 -> destination pitch 3001
    const size_t rows = 5000, rb = 3001, sp = 7000;
    std::vector<char> c(rows * rb, 1);
    host_copy_parallel(c.data(), rb, a.data(), sp, rb, rows);
    for (size_t r = 0; r < rows; ++r)
        if (memcmp(c.data() + r * rb, a.data() + r * sp, rb) != 0) return 3;
    // odd size at odd offsets, nothing written past the end
    std::vector<char> d(8 << 20, 0);
    const size_t len = (5u << 20) + 77;
    host_copy_parallel(d.data() + 3, 0, a.data() + 1, 0, len, 1);
    if (memcmp(d.data() + 3, a.data() + 1, len) != 0 || d[3 + len] != 0 || d[2] != 0) return 4;
    // below the threading threshold
    std::vector<char> e(1000, 0);
    host_copy_parallel(e.data(), 10, a.data(), 20, 10, 100);
    for (size_t r = 0; r < 100; ++r)
        if (memcmp(e.data() + r * 10, a.data() + r * 20, 10) != 0) return 5;
    host_copy_parallel(e.data(), 0, a.data(), 0, 0, 0);  // nothing to do
    const pid_t pid = fork();
    if (pid == 0) {
        std::vector<char> f(n, 0);
        host_copy_parallel(f.data(), n, a.data(), n, n, 1);
        _exit(memcmp(a.data(), f.data(), n) == 0 ? 0 : 1);
    }
    int st = 0;
    waitpid(pid, &st, 0);
    if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) return 6;
    printf("host pool ok (%d threads)\n", host_copy_threads());
    return 0;
}
