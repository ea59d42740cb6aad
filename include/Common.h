// Common.h -- the slice of the reference's include/Common.h that the VPP boundary exposes:
// status codes (reference include/Common.h:19-24), log levels (:29-34) and the Logger handle
// that VideoProcessor::Init receives (:62-70).  Logging / NVTX plumbing itself is out of scope
// for this path; the types exist so that code written against the reference header compiles.
#pragma once
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

enum Internal { VREADER_ERROR = -3, VREADER_UNSUPPORTED = -2, VREADER_REPEAT = -1, VREADER_OK = 0 };
enum LogsLevel { NONE, LOW, MEDIUM, HIGH };

class Logger {
public:
    void initialize(LogsLevel level, std::string logName = "logs.txt") {
        logsLevel = level;
        logFileName = std::move(logName);
        if (!logsFile.is_open() && level > 0) logsFile.open(logFileName);
    }
    ~Logger() {
        if (logsFile.is_open()) logsFile.close();
    }
    std::string logFileName;
    std::ofstream logsFile;
    LogsLevel logsLevel = LogsLevel::NONE;
    bool enableNVTX = false;
};

// slot allocator keyed by consumer name: first use binds a free ("empty") slot forever
// (reference include/Common.h:225-237)
template <class T>
T findFree(const std::string &consumerName, std::vector<std::pair<std::string, T>> &entities) {
    for (auto &item : entities) {
        if (item.first == consumerName) return item.second;
        if (item.first == "empty") {
            item.first = consumerName;
            return item.second;
        }
    }
    return nullptr;
}
