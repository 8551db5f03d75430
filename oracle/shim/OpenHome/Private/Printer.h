// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Printer.h surface (Log::Print to stderr).
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Buffer.h>
#include <cstdio>
#include <cstdarg>

namespace OpenHome {

class Log
{
public:
    static TInt Print(const TChar* aFormat, ...)
    {
        va_list args;
        va_start(args, aFormat);
        const int n = std::vfprintf(stderr, aFormat, args);
        va_end(args);
        return n;
    }
    static TInt Print(const Brx& aMessage)
    {
        return (TInt)std::fwrite(aMessage.Ptr(), 1, aMessage.Bytes(), stderr);
    }
};

} // namespace OpenHome
