// luxrays/utils/properties.h -- the small subset of luxrays::Properties the intersection path
// reads (reference: include/luxrays/utils/properties.h, 754 lines; out of scope beyond key lookup,
// SURVEY.md 2.1).  Supported: Property("key")(default).Get<T>(), Properties << Property(...),
// Properties::Get(Property default), IsDefined, Set, SetFromString ("a.b = v1 v2").
#ifndef _LUXRAYS_B200_PROPERTIES_H
#define _LUXRAYS_B200_PROPERTIES_H

#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace luxrays {

class Property {
public:
	Property() { }
	explicit Property(const std::string &propName) : name(propName) { }

	const std::string &GetName() const { return name; }
	size_t GetSize() const { return values.size(); }

	template <class T> Property &operator()(const T &v) { values.push_back(ToStr(v)); return *this; }
	template <class T> Property &Add(const T &v) { return (*this)(v); }
	Property &Clear() { values.clear(); return *this; }

	template <class T> T Get(const size_t index = 0) const {
		if (index >= values.size())
			throw std::runtime_error("Out of bound error for property: " + name);
		return FromStr<T>(values[index]);
	}
	std::string GetValuesString() const {
		std::string s;
		for (size_t i = 0; i < values.size(); ++i)
			s += (i ? " " : "") + values[i];
		return s;
	}

private:
	template <class T> static std::string ToStr(const T &v) { std::ostringstream ss; ss << v; return ss.str(); }
	static std::string ToStr(const bool &v) { return v ? "1" : "0"; }
	static std::string ToStr(const char *const &v) { return std::string(v); }
	template <class T> static T FromStr(const std::string &s);

	std::string name;
	std::vector<std::string> values;
};

template <> inline std::string Property::FromStr<std::string>(const std::string &s) { return s; }
template <> inline bool Property::FromStr<bool>(const std::string &s) {
	return !(s == "0" || s == "false" || s == "False" || s.empty());
}
template <class T> inline T Property::FromStr(const std::string &s) {
	std::istringstream ss(s);
	T v = T();
	ss >> v;
	if (ss.fail())
		throw std::runtime_error("Unable to parse property value: " + s);
	return v;
}

class Properties {
public:
	Properties() { }

	Properties &Set(const Property &p) { props[p.GetName()] = p; return *this; }
	Properties &operator<<(const Property &p) { return Set(p); }
	Properties &Set(const Properties &o) {
		for (std::map<std::string, Property>::const_iterator it = o.props.begin(); it != o.props.end(); ++it)
			props[it->first] = it->second;
		return *this;
	}
	bool IsDefined(const std::string &name) const { return props.find(name) != props.end(); }
	// returns the stored property, or the argument (carrying its default) when undefined
	Property Get(const Property &defaultProp) const {
		std::map<std::string, Property>::const_iterator it = props.find(defaultProp.GetName());
		return it == props.end() ? defaultProp : it->second;
	}
	Property Get(const std::string &name) const {
		std::map<std::string, Property>::const_iterator it = props.find(name);
		if (it == props.end())
			throw std::runtime_error("Undefined property in Properties::Get(): " + name);
		return it->second;
	}
	// "key = v1 v2 ..." lines; '#' starts a comment
	Properties &SetFromString(const std::string &text) {
		std::istringstream in(text);
		std::string line;
		while (std::getline(in, line)) {
			const size_t hash = line.find('#');
			if (hash != std::string::npos) line = line.substr(0, hash);
			const size_t eq = line.find('=');
			if (eq == std::string::npos) continue;
			std::string key = Trim(line.substr(0, eq));
			if (key.empty()) continue;
			Property p(key);
			std::istringstream vals(line.substr(eq + 1));
			std::string tok;
			while (vals >> tok) {
				if (tok.size() >= 2 && tok.front() == '"' && tok.back() == '"') tok = tok.substr(1, tok.size() - 2);
				p(tok);
			}
			Set(p);
		}
		return *this;
	}
	std::vector<std::string> GetAllNames() const {
		std::vector<std::string> n;
		for (std::map<std::string, Property>::const_iterator it = props.begin(); it != props.end(); ++it) n.push_back(it->first);
		return n;
	}

private:
	static std::string Trim(const std::string &s) {
		const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
		return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
	}
	std::map<std::string, Property> props;
};

}   // namespace luxrays

#endif
