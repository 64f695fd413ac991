// I3CLSimServerInProcess.cxx -- see the header.
#include "I3CLSimServerInProcess.h"

#include <numeric>
#include <stdexcept>

struct I3CLSimServerInProcess::Mailbox {
    std::mutex mutex;
    std::condition_variable ready;
    std::deque<I3CLSimStepToPhotonConverter::ConversionResult_t> results;
    std::string failure;   // set when the server failed while this client was waiting
};

namespace {
std::size_t lcm(std::size_t a, std::size_t b) { return a / std::gcd(a, b) * b; }
} // namespace

I3CLSimServerInProcess::I3CLSimServerInProcess(const std::vector<I3CLSimStepToPhotonConverterPtr> &converters)
    : converters_(converters), workgroupSize_(0), maxBunchSize_(0), shutdown_(false)
{
    if (converters_.empty()) throw std::runtime_error("Need at least 1 I3CLSimStepToPhotonConverter");
    // Harmonize bunch sizes (I3CLSimServer.cxx:95-113)
    for (auto &converter : converters_) {
        if (!converter || !converter->IsInitialized()) throw std::runtime_error("All I3CLSimStepToPhotonConverters must be initialized");
        if (workgroupSize_ == 0) workgroupSize_ = converter->GetWorkgroupSize();
        else workgroupSize_ = lcm(workgroupSize_, converter->GetWorkgroupSize());
        if (maxBunchSize_ == 0) {
            maxBunchSize_ = converter->GetMaxNumWorkitems();
        } else {
            const std::size_t newMaxBunchSize = std::min(maxBunchSize_, converter->GetMaxNumWorkitems());
            const std::size_t newMaxBunchSizeWithGranularity = newMaxBunchSize - newMaxBunchSize % workgroupSize_;
            if (newMaxBunchSizeWithGranularity == 0) throw std::runtime_error("maximum bunch sizes are incompatible with kernel work group sizes.");
            maxBunchSize_ = newMaxBunchSizeWithGranularity;
        }
    }
    const unsigned queueDepth = 5; // I3CLSimServer.cxx:125
    for (unsigned i = 0; i < converters_.size(); i++)
        for (unsigned j = 0; j < queueDepth; j++) workerThreads_.emplace_back(&I3CLSimServerInProcess::WorkerThread, this, i);
}

I3CLSimServerInProcess::~I3CLSimServerInProcess()
{
    {
        std::lock_guard<std::mutex> lock(mutex_);
        shutdown_ = true;
    }
    workAvailable_.notify_all();
    for (auto &thread : workerThreads_) thread.join();
}

std::shared_ptr<I3CLSimClientInProcess> I3CLSimServerInProcess::Connect()
{
    return std::shared_ptr<I3CLSimClientInProcess>(new I3CLSimClientInProcess(this, std::make_shared<Mailbox>()));
}

void I3CLSimServerInProcess::Submit(const std::shared_ptr<Mailbox> &from, I3CLSimStepSeriesConstPtr steps, uint32_t externalId)
{
    {
        std::lock_guard<std::mutex> lock(mutex_);
        if (shutdown_) throw std::runtime_error("I3CLSimServerInProcess is shutting down");
        if (!failure_.empty()) throw std::runtime_error("I3CLSimServerInProcess: a converter failed: " + failure_);
        // assign an internal ID for later reply to the client (I3CLSimServer.cxx:169-180)
        uint32_t internalId = 0;
        if (!clients_.empty()) internalId = (--clients_.end())->first + 1;
        if (clients_.find(internalId) != clients_.end()) throw std::runtime_error("Repeated client ID");
        clients_.emplace(internalId, std::make_pair(from, externalId));
        frontend_.push_back(Task{steps, internalId});
    }
    workAvailable_.notify_one();
}

void I3CLSimServerInProcess::WorkerThread(unsigned index)
{
    for (;;) {
        Task task;
        {
            std::unique_lock<std::mutex> lock(mutex_);
            workAvailable_.wait(lock, [&] { return shutdown_ || !frontend_.empty(); });
            if (frontend_.empty()) return; // shutdown
            task = frontend_.front();
            frontend_.pop_front();
        }
        I3CLSimStepToPhotonConverter::ConversionResult_t result;
        try {
            converters_[index]->EnqueueSteps(task.steps, task.internalId);
            // next result, not necessarily from the batch just enqueued (I3CLSimServer.cxx:318-321)
            result = converters_[index]->GetConversionResult();
        } catch (const std::exception &e) {
            // A converter error is fatal in the reference (the worker's exception ends the server process,
            // I3CLSimServer.cxx:310-343): no result is made up.  The server is marked failed, every waiting client is
            // woken, and Submit / GetConversionResult throw from then on.
            Fail(e.what());
            return;
        }
        std::shared_ptr<Mailbox> destination;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto it = clients_.find(result.identifier);
            if (it == clients_.end()) continue; // "Unknown client ID" (I3CLSimServer.cxx:221-224)
            destination = it->second.first;
            result.identifier = it->second.second; // restore the client's own identifier
            clients_.erase(it);
        }
        {
            std::lock_guard<std::mutex> lock(destination->mutex);
            destination->results.push_back(result);
        }
        destination->ready.notify_one();
    }
}

void I3CLSimServerInProcess::Fail(const std::string &what)
{
    std::vector<std::shared_ptr<Mailbox> > waiting;
    {
        std::lock_guard<std::mutex> lock(mutex_);
        if (failure_.empty()) failure_ = what.empty() ? "unknown error" : what;
        for (auto &c : clients_) waiting.push_back(c.second.first);
        clients_.clear();
        frontend_.clear();
    }
    for (auto &box : waiting) {
        {
            std::lock_guard<std::mutex> lock(box->mutex);
            box->failure = failure_;
        }
        box->ready.notify_all();
    }
}

std::string I3CLSimServerInProcess::Failure() const
{
    std::lock_guard<std::mutex> lock(mutex_);
    return failure_;
}

std::map<std::string, double> I3CLSimServerInProcess::GetStatistics() const
{
    std::map<std::string, double> summary;
    for (std::size_t i = 0; i < converters_.size(); ++i) {
        const std::string postfix = (converters_.size() == 1) ? "" : "_" + std::to_string(i);
        for (auto &v : converters_[i]->GetStatistics()) summary[v.first + postfix] = v.second;
    }
    return summary;
}

I3CLSimClientInProcess::I3CLSimClientInProcess(I3CLSimServerInProcess *server, std::shared_ptr<I3CLSimServerInProcess::Mailbox> mailbox)
    : server_(server), mailbox_(mailbox), workgroupSize_(server->GetWorkgroupSize()), maxBunchSize_(server->GetMaxNumWorkitems()), pending_(0)
{
}

void I3CLSimClientInProcess::EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t identifier)
{
    server_->Submit(mailbox_, steps, identifier);
    pending_++;
}

I3CLSimStepToPhotonConverter::ConversionResult_t I3CLSimClientInProcess::GetConversionResult()
{
    I3CLSimStepToPhotonConverter::ConversionResult_t result;
    if (pending_ != 0) { // I3CLSimServer.cxx:394-419
        std::unique_lock<std::mutex> lock(mailbox_->mutex);
        mailbox_->ready.wait(lock, [&] { return !mailbox_->results.empty() || !mailbox_->failure.empty(); });
        if (mailbox_->results.empty()) throw std::runtime_error("I3CLSimServerInProcess: a converter failed: " + mailbox_->failure);
        result = mailbox_->results.front();
        mailbox_->results.pop_front();
        pending_--;
    }
    return result;
}
