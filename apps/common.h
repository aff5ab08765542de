// common.h -- shared by the app drivers: timing and an optional full dump of the result
// (the reference apps print only the first 10-25 vertices; tests want everything).
#ifndef GM_APPS_COMMON_H
#define GM_APPS_COMMON_H
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

static inline double now_ms() {
  struct timeval t;
  gettimeofday(&t, 0);
  return t.tv_sec * 1e3 + t.tv_usec * 1e-3;
}
// --dump <file> anywhere on the command line
static inline const char* dump_path(int argc, char** argv) {
  for (int i = 1; i + 1 < argc; i++)
    if (!strcmp(argv[i], "--dump")) return argv[i + 1];
  return nullptr;
}
#endif
