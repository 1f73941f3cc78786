// Library-level C-ABI pieces: version + last-error string (thread-local).
#include <cstdarg>
#include <cstdio>
#include "common.cuh"
#include "../../include/cb200.h"

static thread_local char g_err[512] = "";

void cb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* cb_last_error(void) { return g_err; }
extern "C" int cb_version(void) { return 100; }

extern "C" int cb_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(cb_edge_feat_args);
        case 1: return (int)sizeof(cb_tp_segment);
        case 2: return (int)sizeof(cb_tp_conv_args);
        case 3: return (int)sizeof(cb_sde_step_args);
        case 4: return (int)sizeof(cb_tp_row);
        case 5: return (int)sizeof(cb_tp_term);
        case 6: return (int)sizeof(cb_tp_run);
        default: return -1;
    }
}
