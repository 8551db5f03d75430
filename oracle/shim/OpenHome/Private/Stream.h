// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Stream.h surface (IWriter and friends).
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Buffer.h>
#include <OpenHome/Exception.h>

namespace OpenHome {

EXCEPTION(ReaderError);
EXCEPTION(WriterError);

class IWriter
{
public:
    virtual void Write(TByte aValue) = 0;
    virtual void Write(const Brx& aBuffer) = 0;
    virtual void WriteFlush() = 0;
    virtual ~IWriter() {}
};

class IReader
{
public:
    virtual Brn Read(TUint aBytes) = 0;
    virtual void ReadFlush() = 0;
    virtual void ReadInterrupt() = 0;
    virtual ~IReader() {}
};

class WriterBuffer : public IWriter
{
public:
    WriterBuffer(Bwx& aBuffer) : iBuffer(aBuffer) {}
    void Flush() { iBuffer.SetBytes(0); }
    void Write(TByte aValue) override { iBuffer.Append(aValue); }
    void Write(const Brx& aBuffer) override { iBuffer.Append(aBuffer); }
    void WriteFlush() override {}
private:
    Bwx& iBuffer;
};

} // namespace OpenHome
