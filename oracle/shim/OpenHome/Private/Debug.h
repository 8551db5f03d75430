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
#define LOG(x, ...) do { } while (0)
#define LOG_DEBUG(x, ...) do { } while (0)
#define LOG_INFO(x, ...) do { } while (0)
#define LOG_WARNING(x, ...) do { } while (0)
#define LOG_ERROR(x, ...) do { OpenHome::Log::Print(__VA_ARGS__); } while (0)
#define LOG_TRACE(x, ...) do { } while (0)
