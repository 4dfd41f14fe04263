// host/ini_edit.h — writing INI files the way the reference's generators do.
//
// `spinwalk dwi` edits a config in place and `spinwalk config` creates two new ones; both go through mINI 0.9.17
// (MINI_CASE_SENSITIVE): INIFile::write(ini, true) and INIFile::generate(ini, true)
// (reference call sites: src/dwi/pgse.cpp:93-95,145, src/config/config_generator.cpp:176-193; mINI itself is the
// reference's vendored third-party header include/ini.h:435-690).  For a drop-in the files must come out byte for byte,
// so the observable rules of those two operations are restated here (from their behaviour, not their code):
//
//  create (generate)   sections in insertion order as "[name]", each followed by its "key = value" lines (pretty) or
//                      "key=value"; '=' inside a key is written "\="; values are trimmed; sections are separated by one
//                      blank line (pretty); lines are joined with '\n' and the file does NOT end with a newline.
//  update (write)      the existing file is kept line by line (comments, blank lines, spelling and spacing of untouched
//                      entries); lines the parser cannot classify and key lines before the first section are dropped;
//                      an entry whose value changed keeps everything up to the start of its old value and gets the new
//                      value (a single space is inserted when '=' was directly followed by the value);
//                      entries removed from the document are dropped; new keys of an existing section are inserted
//                      after that section's last surviving entry line (or its header); a section removed from the document
//                      loses its header, its entries and the first empty line that follows, its comments stay;
//                      new sections are appended at the end, preceded by a blank line when the file does not end with one.
//                      A missing file is created like `create`.  A UTF-8 BOM is preserved.
#pragma once

#include <string>
#include <utility>
#include <vector>

namespace swk_host {

class IniDocument {
public:
    using Entries = std::vector<std::pair<std::string, std::string>>;

    void clear() { sections_.clear(); }
    // names and keys are trimmed; assigning to an existing key keeps its position
    void set(const std::string &section, const std::string &key, const std::string &value);
    void touch_section(const std::string &section); // creates an empty section if absent (mINI: ini[section])
    bool has_section(const std::string &section) const;
    bool has(const std::string &section, const std::string &key) const;
    std::string get(const std::string &section, const std::string &key) const; // "" when absent
    void remove(const std::string &section, const std::string &key);
    void remove_section(const std::string &section);

    // parse `text` (the rules of host/ini_file.h); when lines != nullptr it receives the lines an update keeps
    void parse(const std::string &text, std::vector<std::string> *lines = nullptr, bool *bom = nullptr);
    bool load(const std::string &path, std::vector<std::string> *lines = nullptr, bool *bom = nullptr);

    std::string render(bool pretty) const;                 // `create`
    bool create_file(const std::string &path, bool pretty) const;
    bool update_file(const std::string &path, bool pretty) const; // `update` (falls back to create when the file is missing)
    // the update as a pure function: text of the existing file -> new text
    std::string merged(const std::string &existing_text, bool pretty) const;

private:
    struct Section { std::string name; Entries kv; };
    const Section *find(const std::string &name) const;
    Section *find(const std::string &name);
    std::vector<Section> sections_;
};

} // namespace swk_host
