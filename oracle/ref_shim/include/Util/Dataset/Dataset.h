/* TEST INFRASTRUCTURE (oracle/_ref build only).  The reference's Util/Dataset/Dataset.h sits on liblmdb and protobuf, neither
 * of which is installed; this in-memory table keeps its interface (getRecord, batchAppend, getRecordsCount; Dataset.h:25-38) so
 * the collectors compile and run unmodified.  The LMDB file format is covered by tests/test_lmdb*.py, not here. */
#pragma once
#include <gsl/span>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <utility>
namespace DeepestScatter {
class Dataset {
public:
    struct Settings {
        Settings(std::string path) : path(std::move(path)) {}
        std::string path;
    };
    explicit Dataset(std::shared_ptr<Settings>) {}
    template <class T> size_t getRecordsCount() { return table<T>().size(); }
    template <class T> T getRecord(int32_t recordId) { return table<T>().at(recordId); }
    template <class T> void batchAppend(const gsl::span<T>& examples, int32_t startId)
    {
        auto& t = table<T>();
        for (const auto& e : examples) t[startId++] = e;
    }
    template <class T> std::map<int32_t, T>& table()
    {
        auto& slot = tables[std::type_index(typeid(T))];
        if (!slot) slot = std::make_shared<std::map<int32_t, T>>();
        return *std::static_pointer_cast<std::map<int32_t, T>>(slot);
    }

private:
    std::map<std::type_index, std::shared_ptr<void>> tables;
};
} // namespace DeepestScatter
