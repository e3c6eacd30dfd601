// Runtime singleton, logging and the error path.
#include "common.h"

namespace qb {

Runtime &rt() {
  static Runtime r;
  return r;
}

void fatal(const char *file, int line, const char *func, const char *fmt, ...) {
  Runtime &r = rt();
  FILE *o = r.out ? r.out : stderr;
  fprintf(o, "%sERROR: ", r.prefix);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(o, fmt, ap);
  va_end(ap);
  fprintf(o, " (rank %d, %s:%d in %s())\n", r.rank, file, line, func);
  fflush(o);
  // same behaviour as the reference's errorQuda -> comm_abort: the process ends
  exit(1);
}

void log_msg(int level, const char *fmt, ...) {
  Runtime &r = rt();
  if (r.verbosity < level || r.rank != 0) return;
  FILE *o = r.out ? r.out : stdout;
  fprintf(o, "%s", r.prefix);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(o, fmt, ap);
  va_end(ap);
  fflush(o);
}

}  // namespace qb
