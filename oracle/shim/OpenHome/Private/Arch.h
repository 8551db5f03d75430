// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Arch.h surface (x86-64 little-endian host).
#pragma once
#include <OpenHome/Types.h>

namespace OpenHome {

class Arch
{
public:
    static inline TUint16 BigEndian2(TUint16 x) { return (TUint16)((x >> 8) | (x << 8)); }
    static inline TUint32 BigEndian4(TUint32 x) { return __builtin_bswap32(x); }
    static inline TUint64 BigEndian8(TUint64 x) { return __builtin_bswap64(x); }
    static inline TUint16 LittleEndian2(TUint16 x) { return x; }
    static inline TUint32 LittleEndian4(TUint32 x) { return x; }
    static inline TUint64 LittleEndian8(TUint64 x) { return x; }
};

} // namespace OpenHome
