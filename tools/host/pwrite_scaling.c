/* How does pwrite() into ONE new file scale with threads?  (the .cfrk writer's second phase)
 * usage: pwrite_scaling <dir> <MB per thread>   build: gcc -O2 -pthread */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>

static int fd; static size_t per; static char *src; static size_t chunk;
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void *work(void *a)
{
    size_t t = (size_t)a, done = 0;
    while (done < per) {
        size_t n = per - done < chunk ? per - done : chunk;
        ssize_t w = pwrite(fd, src + (done % (64u << 20)), n, (off_t)(t * per + done));
        if (w <= 0) { perror("pwrite"); exit(1); }
        done += (size_t)w;
    }
    return NULL;
}
static void *mwork(void *a)   /* the same through a shared mapping */
{
    size_t t = (size_t)a;
    memcpy(src + 0, src + 0, 0);
    return (void *)t;
}
int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "/dev/shm";
    size_t mb = argc > 2 ? (size_t)atol(argv[2]) : 64;
    per = mb << 20;
    src = malloc(64u << 20);
    memset(src, 'x', 64u << 20);
    char path[512];
    snprintf(path, sizeof path, "%s/pwrite_scaling.tmp", dir);
    printf("cores online: %ld\n", sysconf(_SC_NPROCESSORS_ONLN));
    size_t chunks[2] = {per, 4u << 20};
    for (int c = 0; c < 2; c++) {
        chunk = chunks[c];
        for (int n = 1; n <= 32; n *= 2) {
            unlink(path);
            fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
            pthread_t th[32];
            double t0 = now();
            for (size_t i = 0; i < (size_t)n; i++) pthread_create(&th[i], NULL, work, (void *)i);
            for (int i = 0; i < n; i++) pthread_join(th[i], NULL);
            double dt = now() - t0;
            close(fd);
            printf("pwrite chunk %4zu MB, %2d threads x %zu MB: %7.1f ms  %.2f GB/s\n", chunk >> 20, n, mb, dt * 1e3, n * (double)per / dt / 1e9);
        }
    }
    /* mmap + memcpy into a fresh file of the final size */
    for (int n = 1; n <= 32; n *= 4) {
        unlink(path);
        fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (ftruncate(fd, (off_t)(n * per))) { perror("ftruncate"); return 1; }
        char *m = mmap(NULL, n * per, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) { perror("mmap"); return 1; }
        double t0 = now();
        #pragma omp parallel for num_threads(n)
        for (int i = 0; i < n; i++) memcpy(m + (size_t)i * per, src, per < (64u << 20) ? per : (64u << 20));
        double dt = now() - t0;
        munmap(m, n * per); close(fd);
        printf("mmap memcpy, %2d threads x %zu MB: %7.1f ms  %.2f GB/s\n", n, mb, dt * 1e3, n * (double)per / dt / 1e9);
    }
    unlink(path);
    (void)mwork;
    return 0;
}
