// FrameRing.h -- the buffer side of the reference's Decoder (src/Decoder.cpp:97-131 GetFrame, :133-160 the tail
// of Decode, :84-93 notifyConsumers) without the FFmpeg / NVDEC plumbing: a ring of the last `bufferDeep` decoded
// surfaces, per-consumer "new frame" flags, and GetFrame(index) with the reference's index semantics.
// What publishes into it (NVDEC through FFmpeg in the reference, a synthetic producer in tests/cpp/ring_main.cpp) is
// outside the VPP path; what reads from it is the unchanged TensorStream::getFrame sequence
//     decoder->GetFrame(index, consumer, frame);  vpp->Convert(frame, out, params, consumer);
// (src/Wrappers/WrapperPython.cpp:296-312), whose Convert calls this library coalesces into batched launches.
#pragma once
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Common.h"
#include "tsvpp_avframe.h"

class FrameRing {
public:
    explicit FrameRing(int bufferDeep = 10) : deep(bufferDeep), frames(bufferDeep), valid(bufferDeep, false) {}

    // what Decoder::Decode does once avcodec_receive_frame has produced a surface (src/Decoder.cpp:147-159):
    // the slot currentFrame % bufferDeep is replaced, currentFrame advances, every consumer may take it
    void Publish(const AVFrame &decoded) {
        std::unique_lock<std::mutex> locker(sync);
        const int slot = (int)(currentFrame % (unsigned)deep);
        if (valid[slot]) av_frame_unref(&frames[slot]);
        av_frame_ref(&frames[slot], &decoded);
        valid[slot] = true;
        currentFrame++;
        published.store(currentFrame, std::memory_order_release);
        for (auto &item : consumerStatus) item.second = true;
        if (sleepers) consumerSync.notify_all();  // nobody to wake at high rates: consumers that spin see `published`
    }

    // reference src/Decoder.cpp:97-131, statement for statement: blocks until a frame newer than the consumer's
    // last one exists; index > 0 is clamped to 0; index in [-bufferDeep+1, 0] selects an older surface; the slot
    // arithmetic does NOT wrap ((currentFrame-1) % deep + index < 0 -> VREADER_REPEAT), exactly like the reference.
    // Returns currentFrame (the 1-based number of the newest frame), VREADER_REPEAT, or throws when finished.
    int GetFrame(int index, const std::string &consumerName, AVFrame *outputFrame) {
        std::unique_lock<std::mutex> locker(sync);
        if (consumerStatus.find(consumerName) == consumerStatus.end()) {
            consumerStatus[consumerName] = false;
            // frames were published already: hand out the last one without waiting for the next
            if (currentFrame > 0) consumerStatus[consumerName] = true;
        }
        if (!isFinished && !consumerStatus[consumerName]) {
            // a consumer that is faster than the source first polls the frame counter for a few microseconds (a
            // futex sleep and wake-up cost more than a frame period at 1e5+ frames/s), then blocks like the reference
            const unsigned seen = currentFrame;
            locker.unlock();
            for (int spin = 0; spin < 4000 && published.load(std::memory_order_acquire) == seen && !finishedFlag.load(); ++spin) {
            }
            locker.lock();
            ++sleepers;
            while (!consumerStatus[consumerName] && !isFinished) consumerSync.wait(locker);
            --sleepers;
        }
        if (isFinished) throw std::runtime_error("Decoding finished");
        consumerStatus[consumerName] = false;
        if (index > 0) index = 0;
        const int allignedIndex = (int)((currentFrame - 1) % (unsigned)deep) + index;
        if (allignedIndex < 0 || !valid[allignedIndex]) return VREADER_REPEAT;
        av_frame_ref(outputFrame, &frames[allignedIndex]);
        return (int)currentFrame;
    }

    // src/Decoder.cpp:84-93
    int notifyConsumers() {
        std::unique_lock<std::mutex> locker(sync);
        for (auto &item : consumerStatus) item.second = true;
        isFinished = true;
        finishedFlag.store(true);
        consumerSync.notify_all();
        return VREADER_OK;
    }

    unsigned int getFrameIndex() {
        std::unique_lock<std::mutex> locker(sync);
        return currentFrame;
    }

private:
    int deep;
    std::vector<AVFrame> frames;
    std::vector<bool> valid;
    unsigned int currentFrame = 0;
    std::map<std::string, bool> consumerStatus;
    std::mutex sync;
    std::condition_variable consumerSync;
    bool isFinished = false;
    std::atomic<unsigned> published{0};   // = currentFrame, readable without the lock
    std::atomic<bool> finishedFlag{false};
    int sleepers = 0;                     // consumers inside consumerSync.wait (guarded by `sync`)
};
