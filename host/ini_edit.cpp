#include "ini_edit.h"

#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace swk_host {

namespace {
const char *const kWs = " \t\n\r\f\v";

std::string trimmed(const std::string &s)
{
    const size_t b = s.find_first_not_of(kWs);
    if (b == std::string::npos) return "";
    return s.substr(b, s.find_last_not_of(kWs) - b + 1);
}

// position of the first '=' that is not written "\=", or npos
size_t split_at(const std::string &line)
{
    for (size_t i = 0; i < line.size(); i++) {
        if (line[i] == '\\' && i + 1 < line.size() && line[i + 1] == '=') { i++; continue; }
        if (line[i] == '=') return i;
    }
    return std::string::npos;
}

std::string escaped_key(std::string key)
{
    for (size_t p = 0; (p = key.find('=', p)) != std::string::npos; p += 2) key.replace(p, 1, "\\=");
    return key;
}

enum class Kind { Blank, Comment, Header, Entry, Junk };
struct Parsed { Kind kind = Kind::Junk; std::string a, b; };

Parsed classify(const std::string &raw)
{
    Parsed p;
    std::string line = trimmed(raw);
    if (line.empty()) { p.kind = Kind::Blank; return p; }
    if (line[0] == ';') { p.kind = Kind::Comment; return p; }
    if (line[0] == '[') {
        line = line.substr(0, line.find(';'));
        const size_t close = line.rfind(']');
        if (close != std::string::npos) {
            p.kind = Kind::Header;
            p.a = trimmed(line.substr(1, close - 1));
            return p;
        }
    }
    const size_t eq = split_at(line);
    if (eq == std::string::npos) return p; // Junk
    p.kind = Kind::Entry;
    p.a = trimmed(line.substr(0, eq));
    for (size_t q = 0; (q = p.a.find("\\=", q)) != std::string::npos; q += 1) p.a.replace(q, 2, "=");
    p.b = trimmed(line.substr(eq + 1));
    return p;
}

std::string entry_line(const std::string &key, const std::string &value, bool pretty)
{
    return escaped_key(key) + (pretty ? " = " : "=") + trimmed(value);
}

std::string joined(const std::vector<std::string> &lines)
{
    std::string out;
    for (size_t i = 0; i < lines.size(); i++) {
        if (i) out += '\n';
        out += lines[i];
    }
    return out;
}
} // namespace

const IniDocument::Section *IniDocument::find(const std::string &name) const
{
    const std::string n = trimmed(name);
    for (const Section &s : sections_)
        if (s.name == n) return &s;
    return nullptr;
}
IniDocument::Section *IniDocument::find(const std::string &name) { return const_cast<Section *>(static_cast<const IniDocument *>(this)->find(name)); }

void IniDocument::touch_section(const std::string &section)
{
    if (!find(section)) sections_.push_back({trimmed(section), {}});
}

void IniDocument::set(const std::string &section, const std::string &key, const std::string &value)
{
    touch_section(section);
    Section *s = find(section);
    const std::string k = trimmed(key);
    for (auto &kv : s->kv)
        if (kv.first == k) { kv.second = value; return; }
    s->kv.emplace_back(k, value);
}

bool IniDocument::has_section(const std::string &section) const { return find(section) != nullptr; }

bool IniDocument::has(const std::string &section, const std::string &key) const
{
    const Section *s = find(section);
    if (!s) return false;
    const std::string k = trimmed(key);
    for (const auto &kv : s->kv)
        if (kv.first == k) return true;
    return false;
}

std::string IniDocument::get(const std::string &section, const std::string &key) const
{
    const Section *s = find(section);
    if (!s) return "";
    const std::string k = trimmed(key);
    for (const auto &kv : s->kv)
        if (kv.first == k) return kv.second;
    return "";
}

void IniDocument::remove(const std::string &section, const std::string &key)
{
    Section *s = find(section);
    if (!s) return;
    const std::string k = trimmed(key);
    for (size_t i = 0; i < s->kv.size(); i++)
        if (s->kv[i].first == k) { s->kv.erase(s->kv.begin() + i); return; }
}

void IniDocument::remove_section(const std::string &section)
{
    const std::string n = trimmed(section);
    for (size_t i = 0; i < sections_.size(); i++)
        if (sections_[i].name == n) { sections_.erase(sections_.begin() + i); return; }
}

void IniDocument::parse(const std::string &text, std::vector<std::string> *lines, bool *bom)
{
    sections_.clear();
    if (lines) lines->clear();
    size_t pos = 0;
    const bool has_bom = text.size() >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF;
    if (bom) *bom = has_bom;
    if (has_bom) pos = 3;
    if (text.empty()) return; // an empty file has no lines at all (not one empty line)
    std::string current;
    bool in_section = false;
    std::string raw;
    auto take = [&](const std::string &line) {
        const Parsed p = classify(line);
        if (p.kind == Kind::Header) {
            in_section = true;
            current = p.a;
            touch_section(current);
        } else if (p.kind == Kind::Entry && in_section)
            set(current, p.a, p.b);
        if (lines && p.kind != Kind::Junk && !(p.kind == Kind::Entry && !in_section)) lines->push_back(line);
    };
    for (; pos < text.size(); pos++) {
        const char c = text[pos];
        if (c == '\n') { take(raw); raw.clear(); continue; }
        if (c != '\0' && c != '\r') raw += c;
    }
    take(raw);
}

bool IniDocument::load(const std::string &path, std::vector<std::string> *lines, bool *bom)
{
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f.is_open()) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    parse(ss.str(), lines, bom);
    return true;
}

std::string IniDocument::render(bool pretty) const
{
    std::string out;
    for (size_t i = 0; i < sections_.size(); i++) {
        if (i) out += pretty ? "\n\n" : "\n";
        out += "[" + sections_[i].name + "]";
        for (const auto &kv : sections_[i].kv) out += "\n" + entry_line(kv.first, kv.second, pretty);
    }
    return out;
}

std::string IniDocument::merged(const std::string &existing_text, bool pretty) const
{
    IniDocument before;
    std::vector<std::string> lines;
    bool bom = false;
    before.parse(existing_text, &lines, &bom);

    std::vector<std::string> out;
    std::string current;          // section of the line being looked at
    bool keeping = false;         // inside a section the document still has
    bool dropping = false;        // inside a section the document no longer has
    bool eat_one_blank = false;   // ... whose first following empty line goes too
    size_t anchor = 0;            // where new keys of `current` would be inserted

    auto flush_new_keys = [&]() {
        const Section *now = find(current), *old = before.find(current);
        if (!now || !old) return;
        std::vector<std::string> add;
        for (const auto &kv : now->kv)
            if (!before.has(current, kv.first)) add.push_back(entry_line(kv.first, kv.second, pretty));
        out.insert(out.begin() + anchor, add.begin(), add.end());
    };

    for (size_t i = 0; i < lines.size(); i++) {
        const std::string &line = lines[i];
        const Parsed p = classify(line);
        if (p.kind == Kind::Header) {
            if (keeping) { flush_new_keys(); keeping = false; }
            current = p.a;
            if (has_section(current)) {
                keeping = true;
                dropping = false;
                eat_one_blank = false;
                out.push_back(line);
                anchor = out.size();
            } else {
                dropping = true;
                eat_one_blank = true;
            }
        } else if (p.kind == Kind::Entry) {
            if (!dropping && has(current, p.a)) {
                const std::string value = get(current, p.a);
                if (value == p.b) out.push_back(line);
                else {
                    std::string norm = line;
                    for (size_t q = 0; (q = norm.find("\\=", q)) != std::string::npos; q += 2) norm.replace(q, 2, "  ");
                    const size_t eq = norm.find('=');
                    const size_t val = norm.find_first_not_of(kWs, eq + 1);
                    std::string edited = line.substr(0, val);
                    if (pretty && eq + 1 == val) edited += " ";
                    out.push_back(edited + trimmed(value));
                }
                anchor = out.size();
            }
        } else { // blank or comment
            if (eat_one_blank && line.empty()) eat_one_blank = false;
            else out.push_back(line);
        }
        if (i + 1 == lines.size()) flush_new_keys();
    }
    for (const Section &s : sections_) {
        if (before.has_section(s.name)) continue;
        if (pretty && !out.empty() && !out.back().empty()) out.emplace_back();
        out.push_back("[" + s.name + "]");
        for (const auto &kv : s.kv) out.push_back(entry_line(kv.first, kv.second, pretty));
    }
    return (bom ? std::string("\xEF\xBB\xBF") : std::string()) + joined(out);
}

bool IniDocument::create_file(const std::string &path, bool pretty) const
{
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) return false;
    f << render(pretty);
    return bool(f);
}

bool IniDocument::update_file(const std::string &path, bool pretty) const
{
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return create_file(path, pretty);
    std::ifstream in(path, std::ios::in | std::ios::binary);
    if (!in.is_open()) return false;
    std::ostringstream ss;
    ss << in.rdbuf();
    in.close();
    const std::string text = merged(ss.str(), pretty);
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) return false;
    f << text;
    return bool(f);
}

} // namespace swk_host
