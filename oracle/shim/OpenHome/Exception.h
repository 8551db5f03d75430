// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Exception.h surface.
// EXCEPTION(X) declares a typed exception, THROW(X) raises it with file/line,
// ASSERT* raise AssertionFailed (the reference's tests rely on
// TEST_THROWS(..., AssertionFailed), e.g. Media/Tests/TestMsg.cpp:833-835).
#pragma once
#include <OpenHome/Types.h>
#include <cstdio>

namespace OpenHome {

class Exception
{
public:
    Exception(const TChar* aMsg, const TChar* aFile, TUint aLine) : iMsg(aMsg), iFile(aFile), iLine(aLine) {}
    virtual ~Exception() {}
    const TChar* Message() const { return iMsg; }
    const TChar* File() const { return iFile; }
    TUint Line() const { return iLine; }
private:
    const TChar* iMsg;
    const TChar* iFile;
    TUint iLine;
};

#define EXCEPTION(name) \
    class name : public OpenHome::Exception \
    { \
    public: \
        name(const TChar* aFile, TUint aLine) : OpenHome::Exception(#name, aFile, aLine) {} \
        name(const TChar* aMsg, const TChar* aFile, TUint aLine) : OpenHome::Exception(aMsg, aFile, aLine) {} \
    }

#define THROW(name) throw name(__FILE__, __LINE__)

EXCEPTION(AssertionFailed);

inline void CallAssertHandler(const TChar* aFile, TUint aLine)
{
    throw AssertionFailed(aFile, aLine);
}

} // namespace OpenHome

#define ASSERT(x) do { if (!(x)) { OpenHome::CallAssertHandler(__FILE__, __LINE__); } } while (0)
#define ASSERTS() OpenHome::CallAssertHandler(__FILE__, __LINE__)
#define ASSERT_VA(x, fmt, ...) do { if (!(x)) { std::fprintf(stderr, fmt, __VA_ARGS__); OpenHome::CallAssertHandler(__FILE__, __LINE__); } } while (0)
#ifdef DEFINE_DEBUG
# define ASSERT_DEBUG(x) ASSERT(x)
#else
# define ASSERT_DEBUG(x) do { } while (0)
#endif
