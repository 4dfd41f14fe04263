#include "ini_file.h"

#include <fstream>
#include <sstream>

namespace swk_host {

namespace {
const char *const kWs = " \t\n\r\f\v";

std::string trimmed(const std::string &s)
{
    const size_t b = s.find_first_not_of(kWs);
    if (b == std::string::npos) return "";
    const size_t e = s.find_last_not_of(kWs);
    return s.substr(b, e - b + 1);
}
} // namespace

bool IniFile::load(const std::string &path)
{
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f.is_open()) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    parse(ss.str());
    return true;
}

void IniFile::parse(const std::string &text)
{
    data_.clear();
    order_.clear();
    size_t pos = 0;
    if (text.size() >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) pos = 3;
    Section *cur = nullptr;
    std::string raw;
    auto handle = [&](const std::string &rawline) {
        std::string line = trimmed(rawline);
        if (line.empty() || line[0] == ';') return;
        if (line[0] == '[') {
            line = line.substr(0, line.find(';')); // npos -> whole line; a '[' line without ']' goes on as key/value WITHOUT its comment
            const std::string &head = line;
            const size_t close = head.rfind(']');
            if (close != std::string::npos) {
                const std::string name = trimmed(head.substr(1, close - 1));
                auto it = data_.find(name);
                if (it == data_.end()) {
                    it = data_.emplace(name, Section{}).first;
                    order_.push_back(name);
                }
                cur = &it->second;
                return;
            }
        }
        // first '=' that is not escaped as "\="
        size_t eq = std::string::npos;
        for (size_t i = 0; i < line.size(); i++) {
            if (line[i] == '\\' && i + 1 < line.size() && line[i + 1] == '=') { i++; continue; }
            if (line[i] == '=') { eq = i; break; }
        }
        if (eq == std::string::npos || !cur) return;
        std::string key = trimmed(line.substr(0, eq));
        for (size_t p = 0; (p = key.find("\\=", p)) != std::string::npos; p += 1) key.replace(p, 2, "=");
        const std::string value = trimmed(line.substr(eq + 1));
        auto it = cur->index.find(key);
        if (it != cur->index.end()) cur->kv[it->second].second = value;
        else {
            cur->index[key] = cur->kv.size();
            cur->kv.emplace_back(key, value);
        }
    };
    for (; pos < text.size(); pos++) {
        const char c = text[pos];
        if (c == '\n') { handle(raw); raw.clear(); continue; }
        if (c != '\0' && c != '\r') raw += c;
    }
    handle(raw);
}

bool IniFile::has_section(const std::string &section) const { return data_.count(trimmed(section)) == 1; }

bool IniFile::has(const std::string &section, const std::string &key) const
{
    auto s = data_.find(trimmed(section));
    return s != data_.end() && s->second.index.count(trimmed(key)) == 1;
}

std::string IniFile::get(const std::string &section, const std::string &key) const
{
    auto s = data_.find(trimmed(section));
    if (s == data_.end()) return "";
    auto k = s->second.index.find(trimmed(key));
    return k == s->second.index.end() ? "" : s->second.kv[k->second].second;
}

std::vector<std::pair<std::string, std::string>> IniFile::items(const std::string &section) const
{
    auto s = data_.find(trimmed(section));
    return s == data_.end() ? std::vector<std::pair<std::string, std::string>>{} : s->second.kv;
}

} // namespace swk_host
