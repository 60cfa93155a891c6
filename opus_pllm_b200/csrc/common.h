// Process-wide error string and launch counter shared by all translation units.
#pragma once
#include "../../include/opus_b200.h"

namespace opus {
// records `what` (and the pending CUDA error, if any) for opus_last_error(); returns rc
int fail(int rc, const char* what);
// counts kernel launches issued by the library (bench.py reports it as gpu_launches)
void note_launch(long long n = 1);
long long launch_count(bool reset);
}  // namespace opus
