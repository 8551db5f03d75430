// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Optional.h surface.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>

namespace OpenHome {

template <class T> class Optional
{
public:
    Optional(T* aPtr) : iPtr(aPtr) {}
    Optional(T& aRef) : iPtr(&aRef) {}
    Optional(std::nullptr_t) : iPtr(nullptr) {}
    Optional() : iPtr(nullptr) {}
    TBool Ok() const { return iPtr != nullptr; }
    T& Unwrap() const { ASSERT(iPtr != nullptr); return *iPtr; }
    T* Ptr() const { return iPtr; }
private:
    T* iPtr;
};

} // namespace OpenHome
