// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/OsWrapper.h surface (unused by the oracle build).
#pragma once
#include <OpenHome/Types.h>
