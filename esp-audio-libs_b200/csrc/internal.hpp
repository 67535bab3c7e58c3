// Internal links between the translation units of the host layer (api.cu, groups.cu).  Not part of the C ABI.
#pragma once
#include "plan.hpp"

struct EspbResampleBatch;

namespace espb {

// A context that only carries the reference's per-context position state (outputOffset, inputIndex) and geometry —
// what espb_resampleAdvancePosition / GetPosition / GetState / GetRequiredSamples / GetExpectedOutput work on — with
// no device memory of its own.  The fused clock groups hand these out per group; processing calls on them fail.
EspbResampleBatch *new_state_only_context(int num_streams, int channels, const ArtGeometry &geo, float lowpass);
ArtState context_state(const EspbResampleBatch *c);
void set_context_state(EspbResampleBatch *c, ArtState st);
int context_mode(const EspbResampleBatch *c);
void set_context_mode(EspbResampleBatch *c, int mode);

int api_fail(int code, const char *what, const char *detail = nullptr);  // records the thread's last error + status
void api_ok();

}  // namespace espb
