// api.cu — version / status / error text, and host-side q-range resolution.
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "common.cuh"

namespace b200q {

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

static thread_local int g_inputs_resident = 0;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool inputs_resident() { return g_inputs_resident != 0; }

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

// python round(): half-to-even on a double that is k or k+0.5
static double round_half_even(double v) { return nearbyint(v); }

bool make_qspec(int qtype, int symmetric, int reduce_range, QSpec* out) {
  // core/_dtypes.py:8-31
  int full[2], sym[2], red[2];
  bool has_sym = false;
  switch (qtype) {
    case B200Q_UINT4: full[0] = 0; full[1] = 15; red[0] = 0; red[1] = 7; break;
    case B200Q_INT4:  full[0] = -8; full[1] = 7; red[0] = -4; red[1] = 3;
                      sym[0] = -7; sym[1] = 7; has_sym = true; break;
    case B200Q_UINT8: full[0] = 0; full[1] = 255; red[0] = 0; red[1] = 127; break;
    case B200Q_INT8:  full[0] = -128; full[1] = 127; red[0] = -64; red[1] = 64;
                      sym[0] = -127; sym[1] = 127; has_sym = true; break;
    default: return false;
  }
  // core/_dtypes.py:61-70: reduce_range wins, then the symmetric table, else full range
  auto pick = [&](bool is_sym, int* lo, int* hi) {
    if (reduce_range) { *lo = red[0]; *hi = red[1]; }
    else if (is_sym && has_sym) { *lo = sym[0]; *hi = sym[1]; }
    else { *lo = full[0]; *hi = full[1]; }
  };
  QSpec q;
  memset(&q, 0, sizeof(q));
  q.symmetric = symmetric ? 1 : 0;
  pick(symmetric != 0, &q.qmin, &q.qmax);
  pick(false, &q.aqmin, &q.aqmax);
  int slo, shi;
  pick(true, &slo, &shi);
  // utils.py:277-285
  double zero = round_half_even((shi + slo) / 2.0);
  double pos = shi - zero, neg = zero - slo;
  q.sym_zero = (int)zero;
  q.sym_levels = pos < neg ? pos : neg;
  q.is_signed = (qtype == B200Q_INT4 || qtype == B200Q_INT8);
  q.bits = (qtype == B200Q_INT4 || qtype == B200Q_UINT4) ? 4 : 8;
  *out = q;
  return true;
}

}  // namespace b200q

extern "C" {

int b200q_version(void) { return B200Q_VERSION; }

const char* b200q_status_string(int status) {
  switch (status) {
    case B200Q_OK: return "ok";
    case B200Q_ERR_INVALID_ARG: return "invalid argument";
    case B200Q_ERR_UNSUPPORTED: return "unsupported configuration";
    case B200Q_ERR_WORKSPACE: return "workspace too small";
    case B200Q_ERR_CUDA: return "CUDA runtime error";
    case B200Q_NOT_POSITIVE_DEFINITE: return "matrix is not positive definite";
    default: return "unknown status";
  }
}

const char* b200q_last_error(void) { return b200q::g_last_error; }

int b200q_assume_inputs_resident(int on) {
  const int prev = b200q::g_inputs_resident;
  b200q::g_inputs_resident = on ? 1 : 0;
  return prev;
}

long long b200q_launch_count(void) { return b200q::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
