// host/ini_file.h — INI text -> section/key/value map, with the parsing rules the reference gets from
// mINI 0.9.17 built with MINI_CASE_SENSITIVE (reference include/ini.h:273-318 parseLine, :320-420 reader,
// CMakeLists.txt:36).  Written from those rules, not from that code:
//   * the file is split at '\n'; '\r' and NUL bytes are dropped; a UTF-8 BOM is skipped;
//   * a line is trimmed of " \t\n\r\f\v"; empty lines and lines whose first character is ';' are skipped
//     ('#' does NOT start a comment);
//   * "[name]" opens a section: text from the first ';' on is dropped, the name runs to the LAST ']' and is trimmed;
//   * otherwise the first '=' that is not written "\=" splits key and value, both trimmed ("\=" in a key becomes "=");
//     text after the value (e.g. "; comment") stays part of the value — numeric readers stop at it;
//   * key/value lines before the first section are ignored; a repeated key overwrites; names are case-sensitive.
#pragma once

#include <map>
#include <string>
#include <vector>

namespace swk_host {

class IniFile {
public:
    // false when the file cannot be opened
    bool load(const std::string &path);
    void parse(const std::string &text);

    bool has_section(const std::string &section) const;
    bool has(const std::string &section, const std::string &key) const;
    // "" when absent (the reference's ini[section][key] yields an empty string for a missing key)
    std::string get(const std::string &section, const std::string &key) const;
    const std::vector<std::string> &sections() const { return order_; }
    std::vector<std::pair<std::string, std::string>> items(const std::string &section) const;

private:
    struct Section {
        std::map<std::string, size_t> index;
        std::vector<std::pair<std::string, std::string>> kv;
    };
    std::map<std::string, Section> data_;
    std::vector<std::string> order_;
};

} // namespace swk_host
