// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/InfoProvider.h surface.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Buffer.h>
#include <OpenHome/Private/Stream.h>
#include <vector>

namespace OpenHome {

class IInfoProvider
{
public:
    virtual void QueryInfo(const Brx& aQuery, IWriter& aWriter) = 0;
    virtual ~IInfoProvider() {}
};

class IInfoAggregator
{
public:
    virtual void Register(IInfoProvider& aProvider, std::vector<Brn>& aSupportedQueries) = 0;
    virtual ~IInfoAggregator() {}
};

} // namespace OpenHome
