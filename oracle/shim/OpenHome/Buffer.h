// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Buffer.h surface
// (Brx/Brn/Bwx/Bwn/Bws/Bwh).  Containers only; no PCM arithmetic lives here.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>
#include <cstring>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>

namespace OpenHome {

EXCEPTION(BufferOverflow);

class Brx
{
public:
    inline TUint Bytes() const { return iBytes; }
    inline const TByte& operator[](TUint aByteIndex) const { ASSERT(aByteIndex < iBytes); return Ptr()[aByteIndex]; }
    inline const TByte& At(TUint aByteIndex) const { ASSERT(aByteIndex < iBytes); return Ptr()[aByteIndex]; }
    TBool Equals(const Brx& aBrx) const
    {
        if (iBytes != aBrx.iBytes) return false;
        if (iBytes == 0) return true;
        return std::memcmp(Ptr(), aBrx.Ptr(), iBytes) == 0;
    }
    inline TBool operator==(const Brx& aBrx) const { return Equals(aBrx); }
    inline TBool operator!=(const Brx& aBrx) const { return !Equals(aBrx); }
    virtual const TByte* Ptr() const = 0;
    inline class Brn Split(TUint aByteIndex) const;
    inline class Brn Split(TUint aByteIndex, TUint aBytes) const;
    static inline const Brx& Empty();
    virtual ~Brx() {}
protected:
    explicit Brx(TUint aBytes) : iBytes(aBytes) {}
    Brx(const Brx& aBrx) : iBytes(aBrx.iBytes) {}
    TUint iBytes;
private:
    Brx& operator=(const Brx&);
};

class Brn : public Brx
{
public:
    Brn() : Brx(0), iPtr(nullptr) {}
    Brn(const Brx& aBrx) : Brx(aBrx.Bytes()), iPtr(aBrx.Ptr()) {}
    Brn(const Brn& aBrn) : Brx(aBrn.iBytes), iPtr(aBrn.iPtr) {}
    Brn(const TByte* aPtr, TUint aBytes) : Brx(aBytes), iPtr(aPtr) {}
    explicit Brn(const TChar* aPtr) : Brx((TUint)std::strlen(aPtr)), iPtr((const TByte*)aPtr) {}
    inline void Set(const Brx& aBrx) { iPtr = aBrx.Ptr(); iBytes = aBrx.Bytes(); }
    inline void Set(const TByte* aPtr, TUint aBytes) { iPtr = aPtr; iBytes = aBytes; }
    inline void Set(const TChar* aStr) { iPtr = (const TByte*)aStr; iBytes = (TUint)std::strlen(aStr); }
    inline Brn& operator=(const Brn& aBrn) { iPtr = aBrn.iPtr; iBytes = aBrn.iBytes; return *this; }
    const TByte* Ptr() const override { return iPtr; }
protected:
    const TByte* iPtr;
};

inline Brn Brx::Split(TUint aByteIndex) const
{
    ASSERT(aByteIndex <= iBytes);
    return Brn(Ptr() + aByteIndex, iBytes - aByteIndex);
}
inline Brn Brx::Split(TUint aByteIndex, TUint aBytes) const
{
    ASSERT(aByteIndex + aBytes <= iBytes);
    return Brn(Ptr() + aByteIndex, aBytes);
}
inline const Brx& Brx::Empty()
{
    static const Brn kEmpty;
    return kEmpty;
}

// heap-allocated, read-only once set (MimeTypeList.h holds one)
class Brh : public Brx
{
public:
    Brh() : Brx(0), iPtr(nullptr) {}
    explicit Brh(const Brx& aBrx) : Brx(0), iPtr(nullptr) { Set(aBrx); }
    explicit Brh(const TChar* aStr) : Brx(0), iPtr(nullptr) { Set(Brn(aStr)); }
    ~Brh() override { std::free(iPtr); }
    void Set(const Brx& aBrx)
    {
        std::free(iPtr);
        iPtr = (TByte*)std::malloc(aBrx.Bytes() ? aBrx.Bytes() : 1);
        if (aBrx.Bytes() > 0) std::memcpy(iPtr, aBrx.Ptr(), aBrx.Bytes());
        iBytes = aBrx.Bytes();
    }
    void Set(const TChar* aStr) { Set(Brn(aStr)); }
    const TByte* Ptr() const override { return iPtr; }
private:
    Brh(const Brh&);
    Brh& operator=(const Brh&);
    TByte* iPtr;
};

class Bwx : public Brx
{
public:
    inline TUint MaxBytes() const { return iMaxBytes; }
    inline TUint BytesRemaining() const { return iMaxBytes - iBytes; }
    using Brx::At;
    inline TByte& At(TUint aByteIndex) { ASSERT(aByteIndex < iBytes); return const_cast<TByte*>(Ptr())[aByteIndex]; } // (the reference's own tests write through it)
    void SetBytes(TUint aBytes) { ASSERT(aBytes <= iMaxBytes); iBytes = aBytes; }
    void Replace(const Brx& aBuf)
    {
        if (aBuf.Bytes() > iMaxBytes) { ASSERTS(); }
        if (aBuf.Bytes() > 0) std::memmove(const_cast<TByte*>(Ptr()), aBuf.Ptr(), aBuf.Bytes());
        iBytes = aBuf.Bytes();
    }
    void ReplaceThrow(const Brx& aBuf)
    {
        if (aBuf.Bytes() > iMaxBytes) { THROW(BufferOverflow); }
        Replace(aBuf);
    }
    void Replace(const TByte* aPtr, TUint aBytes) { Replace(Brn(aPtr, aBytes)); }
    void Replace(const TChar* aStr) { Replace(Brn(aStr)); }
    void Append(const Brx& aB) { Append(aB.Ptr(), aB.Bytes()); }
    void Append(const TChar* aStr) { Append((const TByte*)aStr, (TUint)std::strlen(aStr)); }
    void Append(const TByte* aPtr, TUint aBytes)
    {
        ASSERT(iBytes + aBytes <= iMaxBytes);
        if (aBytes > 0) std::memcpy(const_cast<TByte*>(Ptr()) + iBytes, aPtr, aBytes);
        iBytes += aBytes;
    }
    void Append(TChar aChar) { TByte b = (TByte)aChar; Append(&b, 1); }
    void Append(TByte aByte) { Append(&aByte, 1); }
    TBool TryAppend(const Brx& aB)
    {
        if (iBytes + aB.Bytes() > iMaxBytes) return false;
        Append(aB);
        return true;
    }
    void AppendThrow(const Brx& aB)
    {
        if (iBytes + aB.Bytes() > iMaxBytes) { THROW(BufferOverflow); }
        Append(aB);
    }
    void AppendPrintf(const TChar* aFormatString, ...)
    {
        va_list args;
        va_start(args, aFormatString);
        const TUint room = iMaxBytes - iBytes;
        TChar* dst = (TChar*)const_cast<TByte*>(Ptr()) + iBytes;
        // vsnprintf always writes a terminator, so format into scratch then copy what fits
        TChar scratch[1024];
        int n = std::vsnprintf(scratch, sizeof scratch, aFormatString, args);
        va_end(args);
        if (n < 0) n = 0;
        if ((TUint)n > sizeof scratch - 1) n = (int)(sizeof scratch - 1);
        if ((TUint)n > room) n = (int)room;
        std::memcpy(dst, scratch, (size_t)n);
        iBytes += (TUint)n;
    }
    const TChar* PtrZ() const
    {
        ASSERT(iBytes < iMaxBytes);
        const_cast<TByte*>(Ptr())[iBytes] = 0;
        return (const TChar*)Ptr();
    }
    void Fill(TByte aFillByte) { std::memset(const_cast<TByte*>(Ptr()), aFillByte, iMaxBytes); }
    void FillZ() { Fill(0); }
    TByte& operator[](TUint aByteIndex) { ASSERT(aByteIndex < iBytes); return const_cast<TByte*>(Ptr())[aByteIndex]; }
    const TByte& operator[](TUint aByteIndex) const { ASSERT(aByteIndex < iBytes); return Ptr()[aByteIndex]; }
protected:
    Bwx(TUint aBytes, TUint aMaxBytes) : Brx(aBytes), iMaxBytes(aMaxBytes) {}
    TUint iMaxBytes;
};

class Bwn : public Bwx
{
public:
    Bwn() : Bwx(0, 0), iPtr(nullptr) {}
    Bwn(const TByte* aPtr, TUint aMaxBytes) : Bwx(0, aMaxBytes), iPtr(aPtr) {}
    Bwn(const TByte* aPtr, TUint aBytes, TUint aMaxBytes) : Bwx(aBytes, aMaxBytes), iPtr(aPtr) { ASSERT(aBytes <= aMaxBytes); }
    void Set(const TByte* aPtr, TUint aMaxBytes) { iPtr = aPtr; iBytes = 0; iMaxBytes = aMaxBytes; }
    void Set(const TByte* aPtr, TUint aBytes, TUint aMaxBytes) { iPtr = aPtr; iBytes = aBytes; iMaxBytes = aMaxBytes; }
    const TByte* Ptr() const override { return iPtr; }
protected:
    const TByte* iPtr;
};

template <TUint S> class Bws : public Bwx
{
public:
    Bws() : Bwx(0, S) {}
    explicit Bws(TUint aBytes) : Bwx(aBytes, S) { ASSERT(aBytes <= S); }
    explicit Bws(const TChar* aStr) : Bwx(0, S) { Replace(aStr); }
    Bws(const TByte* aPtr, TUint aBytes) : Bwx(0, S) { Replace(aPtr, aBytes); }
    Bws(const Brx& aBuf) : Bwx(0, S) { Replace(aBuf); }
    Bws(const Bws<S>& aBuf) : Bwx(0, S) { Replace(aBuf); }
    Bws<S>& operator=(const Bws<S>& aBuf) { if (this != &aBuf) Replace(aBuf); return *this; }
    const TByte* Ptr() const override { return iBuf; }
protected:
    TByte iBuf[S];
};

class Bwh : public Bwx
{
public:
    Bwh() : Bwx(0, 0), iPtr(nullptr) {}
    explicit Bwh(TUint aMaxBytes) : Bwx(0, aMaxBytes), iPtr((TByte*)std::malloc(aMaxBytes ? aMaxBytes : 1)) {}
    Bwh(TUint aBytes, TUint aMaxBytes) : Bwx(aBytes, aMaxBytes), iPtr((TByte*)std::malloc(aMaxBytes ? aMaxBytes : 1)) { ASSERT(aBytes <= aMaxBytes); }
    explicit Bwh(const TChar* aStr) : Bwx(0, (TUint)std::strlen(aStr)), iPtr((TByte*)std::malloc(std::strlen(aStr) + 1)) { Replace(aStr); }
    Bwh(const Brx& aBrx) : Bwx(0, aBrx.Bytes()), iPtr((TByte*)std::malloc(aBrx.Bytes() ? aBrx.Bytes() : 1)) { Replace(aBrx); }
    Bwh(const Bwh& aBuf) : Bwx(0, aBuf.Bytes()), iPtr((TByte*)std::malloc(aBuf.Bytes() ? aBuf.Bytes() : 1)) { Replace(aBuf); }
    ~Bwh() override { std::free(iPtr); }
    void Grow(TUint aMaxBytes)
    {
        if (aMaxBytes > iMaxBytes) {
            TByte* p = (TByte*)std::malloc(aMaxBytes);
            if (iBytes > 0) std::memcpy(p, iPtr, iBytes);
            std::free(iPtr);
            iPtr = p;
            iMaxBytes = aMaxBytes;
        }
    }
    void Set(const Brx& aBrx) { Grow(aBrx.Bytes()); Replace(aBrx); }
    const TByte* Ptr() const override { return iPtr; }
private:
    Bwh& operator=(const Bwh&);
    TByte* iPtr;
};

} // namespace OpenHome
