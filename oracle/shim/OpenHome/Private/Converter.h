// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Converter.h -- the four fixed-width readers the WAV / AIFF
// header parsers use (little- and big-endian 16/32-bit values at a byte index of a buffer).
#pragma once
#include <OpenHome/Buffer.h>

namespace OpenHome {

class Converter
{
public:
    static TUint16 LeUint16At(const Brx& aBuf, TUint aIndex) { return (TUint16)(aBuf[aIndex] | (aBuf[aIndex + 1] << 8)); }
    static TUint32 LeUint32At(const Brx& aBuf, TUint aIndex)
    {
        return (TUint32)aBuf[aIndex] | ((TUint32)aBuf[aIndex + 1] << 8) | ((TUint32)aBuf[aIndex + 2] << 16) | ((TUint32)aBuf[aIndex + 3] << 24);
    }
    static TUint16 BeUint16At(const Brx& aBuf, TUint aIndex) { return (TUint16)((aBuf[aIndex] << 8) | aBuf[aIndex + 1]); }
    static TUint32 BeUint32At(const Brx& aBuf, TUint aIndex)
    {
        return ((TUint32)aBuf[aIndex] << 24) | ((TUint32)aBuf[aIndex + 1] << 16) | ((TUint32)aBuf[aIndex + 2] << 8) | (TUint32)aBuf[aIndex + 3];
    }
};

} // namespace OpenHome
