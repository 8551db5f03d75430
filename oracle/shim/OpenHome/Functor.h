// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Functor.h surface, on std::function.
#pragma once
#include <OpenHome/Types.h>
#include <functional>

namespace OpenHome {

class Functor
{
public:
    Functor() {}
    explicit Functor(std::function<void()> aFn) : iFn(aFn) {}
    void operator()() const { if (iFn) iFn(); }
    operator TBool() const { return (TBool)iFn; }
private:
    std::function<void()> iFn;
};

template <class Type> class FunctorGeneric
{
public:
    FunctorGeneric() {}
    explicit FunctorGeneric(std::function<void(Type)> aFn) : iFn(aFn) {}
    void operator()(Type aType) const { if (iFn) iFn(aType); }
    operator TBool() const { return (TBool)iFn; }
private:
    std::function<void(Type)> iFn;
};

template <class Object, class CallType>
inline Functor MakeFunctor(Object& aC, void (CallType::* const &aF)())
{
    Object* obj = &aC;
    auto fn = aF;
    return Functor([obj, fn]() { (obj->*fn)(); });
}

template <class Type, class Object, class CallType>
inline FunctorGeneric<Type> MakeFunctorGeneric(Object& aC, void (CallType::* const &aF)(Type))
{
    Object* obj = &aC;
    auto fn = aF;
    return FunctorGeneric<Type>([obj, fn](Type aT) { (obj->*fn)(aT); });
}

} // namespace OpenHome
