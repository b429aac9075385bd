// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for allscale/api/core/io.h: a FileIOManager
// singleton handing out copyable stream handles backed by std::fstream.
#pragma once
#include <fstream>
#include <memory>
#include <string>

namespace allscale { namespace api { namespace core {

	enum class Mode { Text, Binary };

	struct Entry { std::string name; Mode mode; };

	class OutputStream {
		std::shared_ptr<std::ofstream> file;
	public:
		OutputStream() {}
		explicit OutputStream(const Entry& e) : file(std::make_shared<std::ofstream>(e.name, e.mode == Mode::Binary ? std::ios::binary | std::ios::out : std::ios::out)) {}
		template<typename T> OutputStream& operator<<(const T& v) { (*file) << v; return *this; }
		OutputStream& operator<<(std::ostream& (*f)(std::ostream&)) { f(*file); return *this; }
		template<typename Op> void atomic(const Op& op) { op(*this); }
		template<typename T> OutputStream& write(const T& v) { file->write(reinterpret_cast<const char*>(&v), sizeof(T)); return *this; }
		explicit operator bool() const { return file && file->good(); }
		void shut() { if (file) file->close(); }
	};

	class InputStream {
		std::shared_ptr<std::ifstream> file;
	public:
		InputStream() {}
		explicit InputStream(const Entry& e) : file(std::make_shared<std::ifstream>(e.name, e.mode == Mode::Binary ? std::ios::binary | std::ios::in : std::ios::in)) {}
		template<typename T> InputStream& operator>>(T& v) { (*file) >> v; return *this; }
		template<typename T> T read() { T v; file->read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }
		explicit operator bool() const { return file && file->good(); }
		std::ifstream& raw() { return *file; }
		void shut() { if (file) file->close(); }
	};

	class FileIOManager {
	public:
		static FileIOManager& getInstance() { static FileIOManager m; return m; }
		Entry createEntry(const std::string& name, Mode mode = Mode::Text) { return Entry{ name, mode }; }
		OutputStream openOutputStream(const Entry& e) { return OutputStream(e); }
		InputStream openInputStream(const Entry& e) { return InputStream(e); }
		void close(OutputStream& s) { s.shut(); }
		void close(InputStream& s) { s.shut(); }
	};

}}}
