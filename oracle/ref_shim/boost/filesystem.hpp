// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/filesystem.hpp>; only path,
// create_directories and remove, which is all the reference uses.
#pragma once
#include <string>
#include <cstdio>
#include <sys/stat.h>
#include <sys/types.h>

namespace boost { namespace filesystem {

class path
{
    std::string p_;
public:
    path() {}
    path(const std::string& s) : p_(s) {}
    path(const char* s) : p_(s) {}
    const std::string& string() const { return p_; }
    const char* c_str() const { return p_.c_str(); }
    path parent_path() const
    {
        std::string::size_type pos = p_.find_last_of('/');
        return pos == std::string::npos ? path("") : path(p_.substr(0, pos));
    }
    bool empty() const { return p_.empty(); }
    // Boost (<1.60 semantics as used by the reference): drop the last path element in place, keeping any trailing separator out
    path& remove_filename()
    {
        std::string::size_type pos = p_.find_last_of('/');
        p_ = (pos == std::string::npos) ? std::string("") : p_.substr(0, pos);
        return *this;
    }
};

inline bool create_directories(const path& p)
{
    const std::string& s = p.string();
    if(s.empty()) return false;
    bool made = false;
    for(std::string::size_type pos = 1; pos <= s.size(); ++pos)
    {
        if(pos == s.size() || s[pos] == '/')
        {
            std::string sub = s.substr(0, pos);
            if(::mkdir(sub.c_str(), 0777) == 0) made = true;
        }
    }
    return made;
}

inline bool remove(const path& p) { return std::remove(p.c_str()) == 0; }
inline bool remove(const std::string& p) { return std::remove(p.c_str()) == 0; }

}} // namespace boost::filesystem
