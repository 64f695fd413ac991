// I3CLSimServerInProcess -- the fan-in / fan-out seam of the reference's I3CLSimServer + I3CLSimClient
// (private/clsim/I3CLSimServer.{h,cxx}) without ZeroMQ: many producers (clients) feed bunches of steps to a
// pool of I3CLSimStepToPhotonConverters (one per GPU), results come back to the client that sent the bunch,
// in any order, tagged with the client's own identifier.
//
// What is kept from the reference, line by line:
//   * bunch-size harmonisation over the converters: granularity = LCM of the workgroup sizes, maximum bunch =
//     the smallest GetMaxNumWorkitems rounded down to the granularity, fatal if that is 0 (I3CLSimServer.cxx:95-113);
//   * 5 worker threads per converter ("queueDepth"), each doing EnqueueSteps(bunch, internal id) followed by
//     GetConversionResult() -- "not necessarily from the batch we just enqueued" (:86-91, 290-343);
//   * internal task ids: the server replaces the client's identifier by its own (largest live id + 1, :169-180)
//     and restores it when the result is routed back (:217-233);
//   * a client learns (workgroupSize, maxNumWorkitems) when it connects (:195-198, 372-381), GetConversionResult
//     returns an empty result when nothing is pending (:394-397);
//   * GetStatistics: every converter's map, keys suffixed "_<index>" when there is more than one (:351-364);
//   * shutdown joins every thread (:119-129).
// What is replaced: the ZeroMQ ROUTER/DEALER sockets and the boost portable-binary archive become in-process
// blocking queues carrying shared_ptrs, so one process drives all GPUs of a box; a work item goes to whichever
// worker is idle first, which is the reference's "workers_ queue" policy (:182-190).
#ifndef I3CLSIMSERVERINPROCESS_H_INCLUDED
#define I3CLSIMSERVERINPROCESS_H_INCLUDED

#ifdef CLSIM_CUDA_IN_ICETRAY
#include "clsim/I3CLSimStepToPhotonConverter.h"
#else
#include "clsim_compat.h"
#endif

#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

class I3CLSimClientInProcess;

class I3CLSimServerInProcess {
public:
    explicit I3CLSimServerInProcess(const std::vector<I3CLSimStepToPhotonConverterPtr> &converters);
    ~I3CLSimServerInProcess();
    I3CLSimServerInProcess(const I3CLSimServerInProcess &) = delete;
    I3CLSimServerInProcess &operator=(const I3CLSimServerInProcess &) = delete;

    std::map<std::string, double> GetStatistics() const;
    std::size_t GetWorkgroupSize() const { return workgroupSize_; }
    std::size_t GetMaxNumWorkitems() const { return maxBunchSize_; }
    // empty while every converter works; otherwise the first converter error (the server no longer accepts or answers work)
    std::string Failure() const;

    // the "servus" handshake: a new client bound to this server
    std::shared_ptr<I3CLSimClientInProcess> Connect();

private:
    friend class I3CLSimClientInProcess;
    struct Mailbox; // a client's result queue
    struct Task {
        I3CLSimStepSeriesConstPtr steps;
        uint32_t internalId;
    };
    void Submit(const std::shared_ptr<Mailbox> &from, I3CLSimStepSeriesConstPtr steps, uint32_t externalId);
    void WorkerThread(unsigned index);
    void Fail(const std::string &what);

    std::vector<I3CLSimStepToPhotonConverterPtr> converters_;
    std::size_t workgroupSize_, maxBunchSize_;

    mutable std::mutex mutex_;
    std::condition_variable workAvailable_;
    std::deque<Task> frontend_;                                                      // bunches waiting for an idle worker
    std::map<uint32_t, std::pair<std::shared_ptr<Mailbox>, uint32_t> > clients_;     // internal id -> (origin, external id)
    bool shutdown_;
    std::string failure_;
    std::vector<std::thread> workerThreads_;
};

class I3CLSimClientInProcess {
public:
    void EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t identifier);
    I3CLSimStepToPhotonConverter::ConversionResult_t GetConversionResult();
    std::size_t GetWorkgroupSize() const { return workgroupSize_; }
    std::size_t GetMaxNumWorkitems() const { return maxBunchSize_; }

private:
    friend class I3CLSimServerInProcess;
    I3CLSimClientInProcess(I3CLSimServerInProcess *server, std::shared_ptr<I3CLSimServerInProcess::Mailbox> mailbox);
    I3CLSimServerInProcess *server_;
    std::shared_ptr<I3CLSimServerInProcess::Mailbox> mailbox_;
    std::size_t workgroupSize_, maxBunchSize_;
    uint32_t pending_;
};

#endif // I3CLSIMSERVERINPROCESS_H_INCLUDED
