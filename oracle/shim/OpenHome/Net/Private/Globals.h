// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Net/Private/Globals.h.  gEnv is only referenced
// under TIMESTAMP_LOGGING_ENABLE, which the reference #undef's (Media/Pipeline/Msg.h:22).
#pragma once
#include <OpenHome/Types.h>
