// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Ascii.h surface (WriterAscii subset).
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Buffer.h>
#include <OpenHome/Private/Stream.h>
#include <cstdio>

namespace OpenHome {

class WriterAscii : public IWriter
{
public:
    WriterAscii(IWriter& aWriter) : iWriter(aWriter) {}
    void Write(TByte aValue) override { iWriter.Write(aValue); }
    void Write(const Brx& aBuffer) override { iWriter.Write(aBuffer); }
    void WriteFlush() override { iWriter.WriteFlush(); }
    void WriteSpace() { iWriter.Write((TByte)' '); }
    void WriteNewline() { iWriter.Write((TByte)'\r'); iWriter.Write((TByte)'\n'); }
    void WriteInt(TInt aValue) { TChar b[16]; std::snprintf(b, sizeof b, "%d", aValue); iWriter.Write(Brn(b)); }
    void WriteUint(TUint aValue) { TChar b[16]; std::snprintf(b, sizeof b, "%u", aValue); iWriter.Write(Brn(b)); }
    void WriteInt64(TInt64 aValue) { TChar b[32]; std::snprintf(b, sizeof b, "%lld", (long long)aValue); iWriter.Write(Brn(b)); }
    void WriteUint64(TUint64 aValue) { TChar b[32]; std::snprintf(b, sizeof b, "%llu", (unsigned long long)aValue); iWriter.Write(Brn(b)); }
    void WriteHex(TUint aValue) { TChar b[16]; std::snprintf(b, sizeof b, "%08x", aValue); iWriter.Write(Brn(b)); }
private:
    IWriter& iWriter;
};

} // namespace OpenHome
