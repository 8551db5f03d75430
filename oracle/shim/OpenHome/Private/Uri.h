// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Uri.h -- Codec/Container.h only holds one by value
#pragma once
#include <OpenHome/Buffer.h>

namespace OpenHome {

class Uri
{
public:
    static const TUint kMaxUriBytes = 1024;
    Uri() {}
    explicit Uri(const Brx& aUri) { iUri.Replace(aUri); }
    void Replace(const Brx& aUri) { iUri.Replace(aUri); }
    const Brx& AbsoluteUri() const { return iUri; }
private:
    Bws<kMaxUriBytes> iUri;
};

} // namespace OpenHome
