// TEST INFRASTRUCTURE ONLY: the generated sources under build/gen/ include "../../include/mifgpu.h" like the
// originals under csrc/ do; from there that path lands here.
#include "../../../include/mifgpu.h"
