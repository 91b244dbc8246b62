/* Debug aid: LD_PRELOAD this to get a backtrace of the main thread on SIGUSR1. */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void handler(int sig) {
  void * frames[64];
  int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  (void)sig;
}
__attribute__((constructor)) static void init(void) {
  struct sigaction sa;
  sa.sa_handler = handler;
  sigemptyset(&sa.sa_mask);
  sa.sa_flags = 0;
  sigaction(SIGUSR1, &sa, 0);
}
