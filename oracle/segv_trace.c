/* ORACLE — TEST / DIAGNOSTIC INFRASTRUCTURE ONLY.
 * Loaded (ctypes) by tools/run_pnp_reference.py before it calls the REFERENCE's native gen_proof: on SIGSEGV / SIGBUS /
 * SIGABRT prints the native backtrace of the faulting thread (glibc backtrace, symbol names from the dynamic symbol
 * tables) to stderr, then exits with 128 + signal.  Used to locate the crash of the reference's own code above
 * HEIGHT=4 on sm_100 without a debugger (under cuda-gdb the crash does not reproduce). */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static void on_fault(int sig, siginfo_t* info, void* ctx) {
    (void)ctx;
    void* frames[64];
    char msg[128];
    int n = snprintf(msg, sizeof(msg), "\n[segv_trace] signal %d at address %p — native backtrace:\n", sig, info ? info->si_addr : 0);
    if (write(2, msg, (size_t)n) < 0) _exit(128 + sig);
    int cnt = backtrace(frames, 64);
    backtrace_symbols_fd(frames, cnt, 2);
    _exit(128 + sig);
}

__attribute__((constructor)) static void install(void) {
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = on_fault;
    sa.sa_flags = SA_SIGINFO | SA_RESETHAND;
    sigaction(SIGSEGV, &sa, 0);
    sigaction(SIGBUS, &sa, 0);
    sigaction(SIGABRT, &sa, 0);
}
