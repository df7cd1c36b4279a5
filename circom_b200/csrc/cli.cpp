// circom_cuda_witness <circuit.cb2c> <input.json> <output.wtns>
//
// Same command-line shape as the reference's generated calculator (`<bin> <input.json> <output.wtns>`,
// c_elements/common/main.cpp:336-373), on top of the C ABI only.  input.json is one input object, or an
// array of input objects (a batch: witnesses are computed together on the GPU and written to
// <output>.0.wtns, <output>.1.wtns, ...).  Input handling follows loadJson / qualify_input /
// json2FrElements (main.cpp:126-286): nested objects and arrays of objects give qualified names
// `a.b[i].c`, values are decimal / 0x / 0b / 0o strings or JSON integers, reduced modulo the prime.
#include <algorithm>
#include <cstdio>
#include <dirent.h>
#include <sys/stat.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/circom_b200.h"
#include "u256.h"

namespace {

// ---- a very small JSON reader (objects, arrays, strings, numbers, true/false/null) -------------------
struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    std::string text;  // Number (verbatim) / String
    bool b = false;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;  // insertion order (nlohmann sorts keys; order is irrelevant here)
};

struct JParser {
    const std::string &s;
    size_t p = 0;
    int depth = 0;  // nesting is bounded: an input file is untrusted and the reader recurses
    static constexpr int MAX_DEPTH = 256;
    explicit JParser(const std::string &str) : s(str) {}
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) ++p; }
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("JSON: ") + m + " at offset " + std::to_string(p)); }
    JValue parse() {
        struct Guard {
            int &d;
            explicit Guard(int &x) : d(x) { ++d; }
            ~Guard() { --d; }
        } guard(depth);
        if (depth > MAX_DEPTH) fail("nesting too deep");
        ws();
        if (p >= s.size()) fail("unexpected end");
        JValue v;
        char c = s[p];
        if (c == '{') {
            v.kind = JValue::Object;
            ++p;
            ws();
            if (p < s.size() && s[p] == '}') { ++p; return v; }
            for (;;) {
                ws();
                if (p >= s.size() || s[p] != '"') fail("expected key");
                std::string k = str();
                ws();
                if (p >= s.size() || s[p] != ':') fail("expected ':'");
                ++p;
                v.obj.emplace_back(k, parse());
                ws();
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JValue::Array;
            ++p;
            ws();
            if (p < s.size() && s[p] == ']') { ++p; return v; }
            for (;;) {
                v.arr.push_back(parse());
                ws();
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = JValue::String;
            v.text = str();
        } else if (c == 't' && s.compare(p, 4, "true") == 0) { v.kind = JValue::Bool; v.b = true; p += 4; }
        else if (c == 'f' && s.compare(p, 5, "false") == 0) { v.kind = JValue::Bool; p += 5; }
        else if (c == 'n' && s.compare(p, 4, "null") == 0) { p += 4; }
        else {
            size_t b = p;
            while (p < s.size() && (isdigit((unsigned char)s[p]) || s[p] == '-' || s[p] == '+' || s[p] == '.' || s[p] == 'e' || s[p] == 'E')) ++p;
            if (b == p) fail("unexpected character");
            v.kind = JValue::Number;
            v.text = s.substr(b, p - b);
        }
        return v;
    }
    std::string str() {
        std::string out;
        ++p;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                char e = s[p + 1];
                out.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
                p += 2;
            } else out.push_back(s[p++]);
        }
        if (p >= s.size()) fail("unterminated string");
        ++p;
        return out;
    }
};

// ---- json2FrElements (main.cpp:144-188) -----------------------------------------------------------------
cw::U256 parse_number(const std::string &text_in, bool is_string, const cw::FieldParams &F) {
    std::string t = text_in;
    unsigned base = 10;
    if (is_string && t.size() >= 2 && t[0] == '0') {
        char c = (char)tolower(t[1]);
        if (c == 'x') { base = 16; t = t.substr(2); }
        else if (c == 'b') { base = 2; t = t.substr(2); }
        else if (c == 'o') { base = 8; t = t.substr(2); }
    }
    bool negative = false;
    if (!is_string) {
        // a JSON number of any spelling goes through a double and is printed with fixed precision 0 (main.cpp:170-175):
        // 3.7 is 4, 2^53 + 1 is 2^53, -5 is q - 5 (mpz_init_set_str reads the sign, mpz_fdiv_r folds it, fr.cpp:2805-2811)
        char buf[400];
        snprintf(buf, sizeof(buf), "%.0f", strtod(t.c_str(), nullptr));
        t = buf;
        if (!t.empty() && t[0] == '-') { negative = true; t = t.substr(1); }
    }
    if (t.empty()) throw std::runtime_error("Invalid number in JSON input: " + text_in);
    cw::U256 acc = cw::u256_from_u64(0), b = cw::u256_from_u64(base);
    for (char ch : t) {
        int d = ch >= '0' && ch <= '9' ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : (ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : 99;
        if (d >= (int)base) throw std::runtime_error("Invalid number in JSON input: " + text_in);
        acc = F.addm(F.mulm(acc, b), cw::u256_from_u64((uint64_t)d));
    }
    if (negative) acc = F.subm(cw::u256_from_u64(0), acc);
    return acc;
}

void flatten_values(const JValue &v, std::vector<cw::U256> &out, const cw::FieldParams &F) {
    if (v.kind == JValue::Array) {
        for (const JValue &e : v.arr) flatten_values(e, out, F);
    } else if (v.kind == JValue::String) out.push_back(parse_number(v.text, true, F));
    else if (v.kind == JValue::Number) out.push_back(parse_number(v.text, false, F));
    else throw std::runtime_error("Invalid JSON type");
}

bool contains_object(const JValue &v) {
    if (v.kind == JValue::Object) return true;
    if (v.kind == JValue::Array)
        for (const JValue &e : v.arr)
            if (contains_object(e)) return true;
    return false;
}
void qualify(const std::string &prefix, const JValue &in, std::vector<std::pair<std::string, const JValue *>> &out);
void qualify_list(const std::string &prefix, const JValue &in, std::vector<std::pair<std::string, const JValue *>> &out) {
    if (in.kind == JValue::Array) {
        for (size_t i = 0; i < in.arr.size(); ++i) qualify_list(prefix + "[" + std::to_string(i) + "]", in.arr[i], out);
    } else qualify(prefix, in, out);
}
// qualify_input (main.cpp:221-241)
void qualify(const std::string &prefix, const JValue &in, std::vector<std::pair<std::string, const JValue *>> &out) {
    if (in.kind == JValue::Array) {
        if (!in.arr.empty() && contains_object(in)) qualify_list(prefix, in, out);
        else out.emplace_back(prefix, &in);
    } else if (in.kind == JValue::Object) {
        for (const auto &kv : in.obj) qualify(prefix.empty() ? kv.first : prefix + "." + kv.first, kv.second, out);
    } else out.emplace_back(prefix, &in);
}

#define CK(call)                                                                      \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != CW_OK) throw std::runtime_error(std::string(#call) + ": " + cw_last_error()); \
    } while (0)

}  // namespace

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "Usage: %s <circuit.cb2c> <input.json | directory of *.json> <output.wtns | output directory>\n", argv[0]);
        return 1;
    }
    try {
        cw_circuit *c = nullptr;
        // CW_O0=1 keeps every signal in the witness (the layout of a reference build with --O0)
        // CW_COMPACT=0: one 32-byte slot per value instead of the compact value store (bit plane + shared temporaries)
        const char *cenv = getenv("CW_COMPACT");
        const uint32_t compact = (cenv && cenv[0] == '0') ? 0u : (uint32_t)CW_FLAG_COMPACT;
        CK(cw_circuit_load(argv[1], (getenv("CW_O0") ? CW_FLAG_O0 : 0) | compact, &c));
        int prime_id = 0;
        CK(cw_circuit_prime(c, &prime_id, nullptr));
        cw::FieldParams F = cw::make_field(prime_id);
        auto slurp = [](const std::string &path) {
            std::ifstream f(path);
            if (!f) throw std::runtime_error("cannot open " + path);
            std::stringstream ss;
            ss << f.rdbuf();
            return ss.str();
        };
        // <input.json>: one input object, or an array of them (a batch: outputs <output.wtns>.<i>.wtns); or a DIRECTORY of
        // *.json files, one input each, taken in name order (outputs <output dir>/<name>.wtns)
        std::vector<JValue> roots;
        std::vector<std::string> out_names;
        std::vector<const JValue *> instances;
        bool is_batch = false;
        struct stat sb;
        const bool is_dir = stat(argv[2], &sb) == 0 && S_ISDIR(sb.st_mode);
        if (is_dir) {
            std::vector<std::string> names;
            if (DIR *dp = opendir(argv[2])) {
                while (struct dirent *e = readdir(dp)) {
                    const std::string n = e->d_name;
                    if (n.size() > 5 && n.compare(n.size() - 5, 5, ".json") == 0) names.push_back(n);
                }
                closedir(dp);
            }
            std::sort(names.begin(), names.end());
            roots.reserve(names.size());
            for (const std::string &n : names) {
                roots.push_back(JParser(slurp(std::string(argv[2]) + "/" + n)).parse());
                out_names.push_back(std::string(argv[3]) + "/" + n.substr(0, n.size() - 5) + ".wtns");
            }
            for (const JValue &r : roots) instances.push_back(&r);
            mkdir(argv[3], 0777);
        } else {
            roots.push_back(JParser(slurp(argv[2])).parse());
            is_batch = roots[0].kind == JValue::Array;
            if (is_batch) for (const JValue &e : roots[0].arr) instances.push_back(&e);
            else instances.push_back(&roots[0]);
        }
        if (instances.empty()) throw std::runtime_error("no inputs");
        cw_batch *b = nullptr;
        CK(cw_batch_create(c, (uint32_t)instances.size(), 0, &b));
        for (size_t i = 0; i < instances.size(); ++i) {
            std::vector<std::pair<std::string, const JValue *>> items;
            qualify("", *instances[i], items);
            for (auto &it : items) {
                uint64_t h = cw_fnv1a(it.first.c_str()), size = 0;
                if (cw_get_input_signal_size(c, h, &size) != CW_OK) throw std::runtime_error("Signal " + it.first + " not found");
                std::vector<cw::U256> vals;
                flatten_values(*it.second, vals, F);
                if (vals.size() < size) throw std::runtime_error("Error loading signal " + it.first + ": Not enough values");
                if (vals.size() > size) throw std::runtime_error("Error loading signal " + it.first + ": Too many values");
                for (size_t k = 0; k < vals.size(); ++k) CK(cw_batch_set_input(b, (uint32_t)i, h, (uint32_t)k, vals[k].v));
            }
            uint32_t rem = 0;
            CK(cw_batch_remaining_inputs(b, (uint32_t)i, &rem));
            if (rem) {
                fprintf(stderr, "Not all inputs have been set. Only %u out of %u\n", cw_get_main_input_signal_no(c) - rem, cw_get_main_input_signal_no(c));
                return 1;
            }
        }
        CK(cw_batch_run(b));
        CK(cw_batch_sync(b));
        std::vector<int32_t> status(instances.size());
        CK(cw_batch_status(b, status.data()));
        for (size_t i = 0; i < instances.size(); ++i) {
            if (status[i] != 0) {
                char where[1024] = "";
                if (status[i] > 0) cw_circuit_assert_info(c, (uint32_t)status[i] - 1, where, sizeof(where), nullptr);   // the reference's line
                fprintf(stderr, "%s%sFailed assert (instance %zu, status %d)\n", where, where[0] ? "\n" : "", i, status[i]);
                return 1;
            }
            {   // what the circuit's log() calls print goes to stdout, as the reference binary prints it (log_bucket.rs:104-162)
                size_t n = 0;
                if (cw_batch_log(b, (uint32_t)i, nullptr, 0, &n) == 0 && n) {
                    std::string text(n + 1, '\0');
                    if (cw_batch_log(b, (uint32_t)i, &text[0], n + 1, &n) == 0) fwrite(text.data(), 1, n, stdout);
                }
            }
            std::string out = argv[3];
            if (is_dir) out = out_names[i];
            else if (is_batch) out += "." + std::to_string(i) + ".wtns";
            CK(cw_batch_write_wtns(b, (uint32_t)i, out.c_str()));
        }
        cw_batch_destroy(b);
        cw_circuit_destroy(c);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
