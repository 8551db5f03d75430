// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Thread.h surface
// (Mutex, AutoMutex, Semaphore, AutoSemaphore) on the C++ standard library.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>
#include <OpenHome/Private/Standard.h>
#include <OpenHome/Functor.h>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <thread>

namespace OpenHome {

EXCEPTION(Timeout);
EXCEPTION(ThreadKill);

enum ThreadPriority {
    kPriorityMuchMore = 4, kPriorityMore = 2, kPriorityLess = -2, kPriorityMuchLess = -4,
    kPrioritySystemLowest = 1, kPriorityLowest = 10, kPriorityVeryLow = 20, kPriorityLow = 30,
    kPriorityNormal = 50, kPriorityHigh = 70, kPriorityVeryHigh = 80, kPriorityHighest = 90,
    kPrioritySystemHighest = 100
};

class Mutex : private INonCopyable
{
public:
    Mutex(const TChar* /*aName*/) {}
    void Wait() { iMutex.lock(); }
    void Signal() { iMutex.unlock(); }
private:
    std::mutex iMutex;
};

class AutoMutex : private INonCopyable
{
public:
    AutoMutex(Mutex& aMutex) : iMutex(aMutex) { iMutex.Wait(); }
    ~AutoMutex() { iMutex.Signal(); }
private:
    Mutex& iMutex;
};

class Semaphore : private INonCopyable
{
public:
    Semaphore(const TChar* /*aName*/, TUint aCount) : iCount(aCount) {}
    void Wait()
    {
        std::unique_lock<std::mutex> lock(iMutex);
        iCv.wait(lock, [this] { return iCount > 0; });
        --iCount;
    }
    void Wait(TUint aTimeoutMs)
    {
        if (aTimeoutMs == 0) { Wait(); return; }
        std::unique_lock<std::mutex> lock(iMutex);
        if (!iCv.wait_for(lock, std::chrono::milliseconds(aTimeoutMs), [this] { return iCount > 0; })) {
            THROW(Timeout);
        }
        --iCount;
    }
    TBool Clear()
    {
        std::lock_guard<std::mutex> lock(iMutex);
        const TBool ret = iCount > 0;
        iCount = 0;
        return ret;
    }
    void Signal()
    {
        { std::lock_guard<std::mutex> lock(iMutex); ++iCount; }
        iCv.notify_one();
    }
private:
    std::mutex iMutex;
    std::condition_variable iCv;
    TUint iCount;
};

class AutoSemaphore : private INonCopyable
{
public:
    AutoSemaphore(Semaphore& aSem) : iSem(aSem) { iSem.Wait(); }
    ~AutoSemaphore() { iSem.Signal(); }
private:
    Semaphore& iSem;
};

class AutoSemaphoreSignal : private INonCopyable
{
public:
    AutoSemaphoreSignal(Semaphore& aSem) : iSem(aSem) {}
    ~AutoSemaphoreSignal() { iSem.Signal(); }
private:
    Semaphore& iSem;
};

// ohNet's Thread: Start() launches Run(); Wait() blocks the thread itself until Signal() (a counting semaphore) and
// throws ThreadKill once Kill() was called; the destructor kills and joins.  (StarvationRamper.cpp's RampGenerator runs
// FlywheelRamperManager::Ramp on such a thread.)
class Thread : private INonCopyable
{
public:
    static const TUint kDefaultStackBytes = 32 * 1024;
    static const TChar* CurrentThreadName() { return "oracle"; }
    static void Sleep(TUint aMilliSecs) { std::this_thread::sleep_for(std::chrono::milliseconds(aMilliSecs)); } // (only the reference's tests sleep)
    virtual ~Thread() { Kill(); Join(); }
    void Start() { iThread = std::thread([this] { try { Run(); } catch (ThreadKill&) {} }); }
    void Wait()
    {
        std::unique_lock<std::mutex> lock(iMutex);
        iCv.wait(lock, [this] { return iCount > 0 || iKill; });
        if (iKill) THROW(ThreadKill);
        --iCount;
    }
    TBool TryWait()
    {
        std::lock_guard<std::mutex> lock(iMutex);
        if (iKill) THROW(ThreadKill);
        if (iCount == 0) return false;
        --iCount;
        return true;
    }
    void Signal()
    {
        { std::lock_guard<std::mutex> lock(iMutex); ++iCount; }
        iCv.notify_all();
    }
    void Kill()
    {
        { std::lock_guard<std::mutex> lock(iMutex); iKill = true; }
        iCv.notify_all();
    }
    void CheckForKill() const { if (iKill) THROW(ThreadKill); }
    void Join() { if (iThread.joinable()) iThread.join(); }
protected:
    Thread(const TChar* /*aName*/, TUint /*aPriority*/ = kPriorityNormal, TUint /*aStackBytes*/ = kDefaultStackBytes) {}
    virtual void Run() = 0;
private:
    std::thread iThread;
    std::mutex iMutex;
    std::condition_variable iCv;
    TUint iCount = 0;
    bool iKill = false;
};

class ThreadFunctor : public Thread
{
public:
    ThreadFunctor(const TChar* aName, Functor aFunctor, TUint aPriority = kPriorityNormal, TUint aStackBytes = kDefaultStackBytes)
        : Thread(aName, aPriority, aStackBytes), iFunctor(aFunctor) {}
    ~ThreadFunctor() { Kill(); Join(); }
private:
    void Run() override { iFunctor(); }
    Functor iFunctor;
};

} // namespace OpenHome
