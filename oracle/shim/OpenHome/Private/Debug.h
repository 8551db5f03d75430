// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Debug.h surface.
// LOG(...) compiles to nothing (ohNet's release behaviour with no debug level set);
// LOG_ERROR prints.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Private/Printer.h>

namespace OpenHome {

class Debug
{
public:
    static const TUint kApplication0 = 1u << 0;
};

} // namespace OpenHome

#define kApplication1 1
#define kApplication2 2
#define kApplication3 3
#define kApplication4 4
#define kApplication5 5
#define kApplication6 6
#define kApplication34 34
// plain brace blocks, as in ohNet: some call sites in the reference carry no trailing semicolon (CodecController.cpp:421)
#define LOG(x, ...) { }
#define LOG_DEBUG(x, ...) { }
#define LOG_INFO(x, ...) { }
#define LOG_WARNING(x, ...) { }
#define LOG_ERROR(x, ...) { OpenHome::Log::Print(__VA_ARGS__); }
#define LOG_TRACE(x, ...) { }
