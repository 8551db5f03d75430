// ohNet shim (TEST INFRASTRUCTURE ONLY) -- minimal stand-in for ohNet 1.40.5859's
// OpenHome/Types.h so that the reference's Msg.cpp compiles unmodified from
// /root/reference.  ohNet is an un-vendored binary dependency of ohPipeline
// (projectdata/dependencies.json:3-17) and is absent from this container.
// Nothing here carries any of the hot path's arithmetic: fixed-width typedefs only.
#pragma once
#include <cstdint>
#include <cstddef>

#ifndef DllExport
# define DllExport
#endif
#ifndef DllExportClass
# define DllExportClass
#endif

typedef bool TBool;
typedef char TChar;
typedef uint8_t TByte;
typedef int8_t TInt8;
typedef uint8_t TUint8;
typedef int16_t TInt16;
typedef uint16_t TUint16;
typedef int32_t TInt32;
typedef uint32_t TUint32;
typedef int64_t TInt64;
typedef uint64_t TUint64;
typedef int TInt;
typedef unsigned int TUint;
typedef void TAny;
typedef uint32_t TIpAddress;

static_assert(sizeof(TUint) == 4, "TUint must be 32-bit for parity with ohNet");
static_assert(sizeof(TInt) == 4, "TInt must be 32-bit for parity with ohNet");
