// Report.h — flat JSON run report with the keys upstream tooling reads.
//
// Mirrors the use the cavity benchmark makes of Neon::Report (libNeonCore/src/core/tools/Report.cpp:22-122: a record name,
// addMember(name, scalar | string | vector), write(fileName, appendTimeToFileName) -> "<fileName>[_<time>].json").
// Written with a hand-rolled emitter: the reference pulls rapidjson for this, which the hot path does not need.
#pragma once

#include <chrono>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "Neon/Neon.h"

namespace Neon {

class Report
{
   public:
    Report() = default;
    explicit Report(const std::string& recordName) { addMember("Record Name", recordName); }

    template <typename T>
    void addMember(const std::string& name, const T& value)
    {
        mMembers.emplace_back(name, encode(value));
    }
    void addMember(const std::string& name, const char* value) { mMembers.emplace_back(name, quote(value)); }

    /* nested object from another report's members (Report::addSubdoc) */
    void addSubdoc(const std::string& name, const Report& sub) { mMembers.emplace_back(name, sub.body(2)); }

    void setToken(const std::string& token) { addMember("token", token); }

    /* returns the path written */
    std::string write(const std::string& outputFilename, bool appendTimeToFileName = true) const
    {
        std::string full = outputFilename;
        if (appendTimeToFileName) {
            const auto  now = std::chrono::system_clock::now();
            std::time_t t = std::chrono::system_clock::to_time_t(now);
            const auto  ms = std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000;
            std::tm     tm{};
            localtime_r(&t, &tm);
            std::ostringstream s;
            s << "_" << std::put_time(&tm, "%Y_%m_%d__%H_%M_%S") << "_" << std::setw(3) << std::setfill('0') << ms;
            full += s.str();
        }
        full += ".json";
        std::ofstream out(full);
        if (!out) {
            NeonException e("Report::write");
            e << "cannot open " << full;
            NEON_THROW(e);
        }
        out << body(0) << "\n";
        return full;
    }

    std::string dump() const { return body(0); }

   private:
    static std::string quote(const std::string& s)
    {
        std::string o = "\"";
        for (char c : s) {
            switch (c) {
                case '"': o += "\\\""; break;
                case '\\': o += "\\\\"; break;
                case '\n': o += "\\n"; break;
                case '\t': o += "\\t"; break;
                default: o += c;
            }
        }
        return o + "\"";
    }
    template <typename T>
    static std::string encode(const T& v)
    {
        if constexpr (std::is_same_v<T, bool>) {
            return v ? "true" : "false";
        } else if constexpr (std::is_same_v<T, std::string>) {
            return quote(v);
        } else if constexpr (std::is_floating_point_v<T>) {
            std::ostringstream s;
            s << std::setprecision(17) << v;
            return s.str();
        } else if constexpr (std::is_integral_v<T>) {
            return std::to_string(v);
        } else {
            std::string o = "[";
            bool        first = true;
            for (const auto& e : v) {
                o += first ? "" : ", ";
                o += encode(std::decay_t<decltype(e)>(e));
                first = false;
            }
            return o + "]";
        }
    }
    std::string body(int indent) const
    {
        const std::string pad(indent + 4, ' ');
        std::string       o = "{\n";
        for (size_t i = 0; i < mMembers.size(); ++i) {
            o += pad + quote(mMembers[i].first) + ": " + mMembers[i].second + (i + 1 < mMembers.size() ? ",\n" : "\n");
        }
        return o + std::string(indent, ' ') + "}";
    }
    std::vector<std::pair<std::string, std::string>> mMembers;
};

}  // namespace Neon
