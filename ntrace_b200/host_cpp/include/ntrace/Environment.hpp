// Environment — option registry, `config.conf` reader and command-line overrides with the reference's key names.
// Reference: src/rt/Environment.cpp:880-1185 (file grammar `Group { key value }`, '#' comments, nested groups;
// `-D<Group.key>=<value>` on the command line; the first positional argument names the environment file) and the option
// table of src/rt/AppEnvironment.cpp:39-174.  Keys outside the tracing path are accepted and kept as strings so that the
// reference's shipped config.conf parses unchanged.
#pragma once
#include "ntrace/Base.hpp"
#include <fstream>
#include <map>
#include <sstream>

class Environment
{
public:
    Environment()
    {
        static const char* defaults[][2] = {
            {"App.benchmark", "true"}, {"App.log", "ntrace.log"}, {"App.stats", "stats.log"}, {"App.frameWidth", "1024"}, {"App.frameHeight", "768"},
            {"Benchmark.warmupRepeats", "1"}, {"Benchmark.measureRepeats", "5"}, {"Renderer.samples", "8"}, {"Renderer.sortRays", "true"},
            {"Renderer.cacheDataStructure", "true"}, {"Renderer.numGpus", "1"}, {"Raygen.random", "false"}, {"Raygen.aoRadius", "5.0"}, {"SBVH.alpha", "1.0e-5"}};
        for (size_t i = 0; i < sizeof(defaults) / sizeof(defaults[0]); i++) m_values[defaults[i][0]] = defaults[i][1];
    }

    static Environment* GetSingleton() { static Environment env; return s_override() ? s_override() : &env; }
    static void SetSingleton(Environment* e) { s_override() = e; }

    // the reference's getters return false when the option is missing
    bool GetStringValue(const std::string& name, std::string& value) const { auto it = m_values.find(name); if (it == m_values.end()) return false; value = it->second; return true; }
    bool GetIntValue(const std::string& name, int& value) const { std::string s; if (!GetStringValue(name, s)) return false; value = atoi(s.c_str()); return true; }
    bool GetFloatValue(const std::string& name, float& value) const { std::string s; if (!GetStringValue(name, s)) return false; value = (float)atof(s.c_str()); return true; }
    bool GetBoolValue(const std::string& name, bool& value) const
    {
        std::string s;
        if (!GetStringValue(name, s)) return false;
        for (size_t i = 0; i < s.size(); i++) s[i] = (char)tolower(s[i]);
        if (s == "true" || s == "on" || s == "yes" || s == "1") { value = true; return true; }
        if (s == "false" || s == "off" || s == "no" || s == "0") { value = false; return true; }
        FW::fail("Environment: option %s has a non-boolean value '%s'", name.c_str(), s.c_str());
        return false;
    }
    bool Has(const std::string& name) const { return m_values.count(name) != 0; }
    void Set(const std::string& name, const std::string& value) { m_values[name] = value; }

    void ReadEnvFile(const std::string& path)
    {
        std::ifstream f(path.c_str());
        if (!f) FW::fail("Error: Cannot open environment file '%s'.", path.c_str());
        std::stringstream ss; ss << f.rdbuf();
        ParseEnvString(ss.str(), path);
    }

    void ParseEnvString(const std::string& text, const std::string& origin = "<string>")
    {
        std::vector<std::string> prefix;
        std::istringstream in(text);
        std::string raw;
        int lineno = 0;
        while (std::getline(in, raw)) {
            lineno++;
            std::string line = raw.substr(0, raw.find('#'));
            std::vector<std::string> toks;
            std::string cur;
            for (size_t i = 0; i <= line.size(); i++) {
                char c = i < line.size() ? line[i] : ' ';
                if (c == '{' || c == '}' || isspace((unsigned char)c)) {
                    if (!cur.empty()) { toks.push_back(cur); cur.clear(); }
                    if (c == '{' || c == '}') toks.push_back(std::string(1, c));
                } else cur += c;
            }
            size_t i = 0;
            while (i < toks.size()) {
                const std::string& tok = toks[i];
                if (tok == "}") {
                    if (prefix.empty()) FW::fail("Error: unpaired } in %s (line %d).", origin.c_str(), lineno);
                    prefix.pop_back(); i++;
                } else if (i + 1 < toks.size() && toks[i + 1] == "{") { prefix.push_back(tok); i += 2; }
                else if (tok == "{") FW::fail("Error: group without a name in %s (line %d).", origin.c_str(), lineno);
                else {
                    size_t j = i + 1;
                    std::string val;
                    while (j < toks.size() && toks[j] != "{" && toks[j] != "}") { if (!val.empty()) val += " "; val += toks[j]; j++; }
                    std::string key;
                    for (size_t k = 0; k < prefix.size(); k++) key += prefix[k] + ".";
                    key += tok;
                    if (val.empty()) FW::fail("Error: option %s has no value in %s (line %d).", key.c_str(), origin.c_str(), lineno);
                    Set(key, val);
                    i = j;
                }
            }
        }
        if (!prefix.empty()) FW::fail("Error: unclosed group %s in %s.", prefix.back().c_str(), origin.c_str());
    }

    // [envfile] -D<Group.key>=<value> ...
    bool Parse(int argc, char** argv, const char* defaultEnvFile = NULL)
    {
        std::string envFile = defaultEnvFile ? defaultEnvFile : "";
        std::vector<std::string> rest;
        bool positionalSeen = false;
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            if (a.empty()) continue;
            if (a[0] != '-' && !positionalSeen && rest.empty()) { envFile = a; positionalSeen = true; }
            else rest.push_back(a);
        }
        if (!envFile.empty()) ReadEnvFile(envFile);
        for (size_t i = 0; i < rest.size(); i++) {
            const std::string& a = rest[i];
            size_t eq = a.find('=');
            if (a.compare(0, 2, "-D") != 0 || eq == std::string::npos) FW::fail("Environment: unknown option '%s'", a.c_str());
            Set(a.substr(2, eq - 2), a.substr(eq + 1));
        }
        return true;
    }

private:
    static Environment*& s_override() { static Environment* p = NULL; return p; }
    std::map<std::string, std::string> m_values;
};
