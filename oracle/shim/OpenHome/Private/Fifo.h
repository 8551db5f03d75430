// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Fifo.h surface (FifoLiteDynamic only).
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>
#include <OpenHome/Private/Standard.h>
#include <OpenHome/Private/Thread.h>
#include <vector>

namespace OpenHome {

EXCEPTION(FifoReadError);

template <class T> class FifoLiteDynamic : private INonCopyable
{
public:
    FifoLiteDynamic(TUint aSlots) : iBuf(aSlots), iSlots(aSlots), iSlotsUsed(0), iReadIndex(0), iWriteIndex(0) {}
    TUint Slots() const { return iSlots; }
    TUint SlotsFree() const { return iSlots - iSlotsUsed; }
    TUint SlotsUsed() const { return iSlotsUsed; }
    void Write(T aEntry)
    {
        ASSERT(iSlotsUsed < iSlots);
        iBuf[iWriteIndex] = aEntry;
        iWriteIndex = (iWriteIndex + 1 == iSlots) ? 0 : iWriteIndex + 1;
        ++iSlotsUsed;
    }
    T Read()
    {
        ASSERT(iSlotsUsed > 0);
        T entry = iBuf[iReadIndex];
        iReadIndex = (iReadIndex + 1 == iSlots) ? 0 : iReadIndex + 1;
        --iSlotsUsed;
        return entry;
    }
private:
    std::vector<T> iBuf;
    TUint iSlots;
    TUint iSlotsUsed;
    TUint iReadIndex;
    TUint iWriteIndex;
};

} // namespace OpenHome
