// Text primitives of the libpsim host: the reference's file formats are whitespace-split on
// read and tab-separated on write, with Java's number syntax on both sides
// (reference: Tools.readWords Tools.java:392-401; Integer.parseInt / Double.parseDouble /
// Double.toString call sites in Main.java:499-1353).
//
// Files are slurped whole and tokenised in place (string_view) so a multi-GB .hsa/.hsb pair
// parses at memory speed; writers format into per-thread row buffers.
#pragma once

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace psimhost {

using sv = std::string_view;

// The reference throws java.lang.Exception(message) and prints the message (Main.java:106-109).
struct HostError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Java String.trim() strips every char <= ' ' from both ends.
inline sv jtrim(sv s) {
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    return s.substr(a, b - a);
}

// Java regex \s = [ \t\n\x0B\f\r]
inline bool jspace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// Tools.readWords: trim, then split on runs of whitespace. Appends to `out` (cleared first).
inline void split_words(sv line, std::vector<sv>& out) {
    out.clear();
    line = jtrim(line);
    size_t i = 0, n = line.size();
    while (i < n) {
        size_t j = i;
        while (j < n && !jspace(line[j])) j++;
        // a non-\s control char (<= ' ' but not in \s) can only sit inside a word after trim
        out.push_back(line.substr(i, j - i));
        while (j < n && jspace(line[j])) j++;
        i = j;
    }
}

inline std::string upper(sv s) {
    std::string r(s);
    for (auto& c : r)
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return r;
}

// Integer.parseInt: optional sign, decimal digits only, must fit int32.
inline bool parse_int32(sv s, int32_t& out) {
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '-' || s[0] == '+') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i == s.size()) return false;
    int64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > (int64_t)1 << 31) return false;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return false;
    out = (int32_t)v;
    return true;
}

// Double.parseDouble on the syntax the formats use: decimal / exponent forms, optional sign,
// "NaN" / "Infinity", optional trailing f/F/d/D. Correctly rounded (strtod).
inline bool parse_double(sv s, double& out) {
    s = jtrim(s);
    if (s.empty() || s.size() > 400) return false;
    if (s.back() == 'f' || s.back() == 'F' || s.back() == 'd' || s.back() == 'D') {
        if (s == "Infinity" || s == "+Infinity" || s == "-Infinity") {
        } else
            s.remove_suffix(1);
        if (s.empty()) return false;
    }
    sv body = s;
    if (body[0] == '+' || body[0] == '-') body.remove_prefix(1);
    if (body == "NaN") {
        out = std::nan("");
        return true;
    }
    if (body == "Infinity") {
        out = s[0] == '-' ? -INFINITY : INFINITY;
        return true;
    }
    if (body.empty()) return false;
    bool digits = false;
    for (char c : body) {  // reject what strtod accepts but Java does not (hex, "inf", "nan(...)")
        if (c >= '0' && c <= '9')
            digits = true;
        else if (c != '.' && c != 'e' && c != 'E' && c != '+' && c != '-')
            return false;
    }
    if (!digits) return false;
    char buf[408];
    std::memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    char* end = nullptr;
    out = std::strtod(buf, &end);
    return end == buf + s.size();
}

// Java Double.toString: the shortest decimal that round-trips (at least one fraction digit);
// plain notation for 1e-3 <= |d| < 1e7, otherwise d.dddE[-]n. Appends to `o`.
inline void java_double(double d, std::string& o) {
    if (std::isnan(d)) {
        o += "NaN";
        return;
    }
    if (std::isinf(d)) {
        o += d < 0 ? "-Infinity" : "Infinity";
        return;
    }
    if (d == 0) {
        o += std::signbit(d) ? "-0.0" : "0.0";
        return;
    }
    char digs[32];
    auto r = std::to_chars(digs, digs + sizeof digs, d, std::chars_format::scientific);  // shortest round-trip
    sv t(digs, (size_t)(r.ptr - digs));
    bool neg = t[0] == '-';
    if (neg) t.remove_prefix(1);
    size_t epos = t.find('e');
    int exp10 = std::atoi(std::string(t.substr(epos + 1)).c_str());
    std::string mant;  // digits without the point
    for (char c : t.substr(0, epos))
        if (c != '.') mant.push_back(c);
    if (neg) o.push_back('-');
    int nd = (int)mant.size();
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            for (int i = 0; i <= exp10; i++) o.push_back(i < nd ? mant[i] : '0');
            o.push_back('.');
            if (nd > exp10 + 1)
                o.append(mant, exp10 + 1, std::string::npos);
            else
                o.push_back('0');
        } else {
            o += "0.";
            for (int i = 0; i < -exp10 - 1; i++) o.push_back('0');
            o += mant;
        }
    } else {
        o.push_back(mant[0]);
        o.push_back('.');
        if (nd > 1)
            o.append(mant, 1, std::string::npos);
        else
            o.push_back('0');
        o.push_back('E');
        o += std::to_string(exp10);
    }
}

// Whole-file reader with a line cursor. Lines end at \n, \r\n or \r like BufferedReader.readLine.
class LineFile {
public:
    bool open(const std::string& path) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) return false;
        std::fseek(f, 0, SEEK_END);
        long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (n < 0) {  // not seekable: stream it
            char tmp[1 << 16];
            size_t k;
            while ((k = std::fread(tmp, 1, sizeof tmp, f)) > 0) data_.append(tmp, k);
        } else {
            data_.resize((size_t)n);
            size_t got = n ? std::fread(&data_[0], 1, (size_t)n, f) : 0;
            data_.resize(got);
        }
        std::fclose(f);
        pos_ = 0;
        return true;
    }
    // next raw line (untrimmed); false at end of file
    bool next(sv& line) {
        if (pos_ >= data_.size()) return false;
        size_t e = pos_;
        const char* p = data_.data();
        size_t n = data_.size();
        while (e < n && p[e] != '\n' && p[e] != '\r') e++;
        line = sv(p + pos_, e - pos_);
        if (e < n && p[e] == '\r' && e + 1 < n && p[e + 1] == '\n') e++;
        pos_ = e + 1;
        return true;
    }
    // next trimmed line that is neither blank nor a ';' comment (the reference's data-line loops)
    bool next_data(sv& line) {
        while (next(line)) {
            line = jtrim(line);
            if (!line.empty() && line[0] != ';') return true;
        }
        return false;
    }

private:
    std::string data_;
    size_t pos_ = 0;
};

inline bool file_readable(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fclose(f);
    return true;
}

// Buffered output file; large rows are appended pre-formatted.
class OutFile {
public:
    explicit OutFile(const std::string& path) : f_(std::fopen(path.c_str(), "wb")) {
        if (!f_) throw HostError(path + " (cannot create file)");
        std::setvbuf(f_, nullptr, _IOFBF, 1 << 22);
    }
    ~OutFile() {
        if (f_) std::fclose(f_);
    }
    void write(const char* p, size_t n) { std::fwrite(p, 1, n, f_); }
    void write(const std::string& s) { std::fwrite(s.data(), 1, s.size(), f_); }
    void close() {
        if (f_) std::fclose(f_);
        f_ = nullptr;
    }

private:
    FILE* f_;
};

}  // namespace psimhost
