// sync_block.h -- minimal stand-in for the GNU Radio 3.7 block base classes: stream signatures,
// work(), consume_each(), message ports with SYNCHRONOUS delivery (a block's handler runs inside
// message_port_pub of the sender; GNU Radio queues it to the receiver's thread -- same order per
// port, which is all gr-amps relies on).  Used only where the real <gnuradio/sync_block.h> is absent.
#pragma once
#include "../boost_standin.h"
#include <gnuradio/attributes.h>
#include <pmt/pmt.h>

#include <complex>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;
typedef std::vector<int> gr_vector_int;


namespace gr {

class io_signature {
public:
    typedef boost::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_item) {
        sptr s(new io_signature());
        s->d_min = min_streams; s->d_max = max_streams; s->d_size = sizeof_item;
        return s;
    }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int) const { return d_size; }
private:
    int d_min = 0, d_max = 0, d_size = 0;
};

class basic_block {
public:
    typedef std::function<void(pmt::pmt_t)> msg_handler_t;
    basic_block() {}   // allows pure virtual interface sub-classes (as in GNU Radio)
    basic_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out) {}
    virtual ~basic_block() {}
    const std::string &name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }
    void message_port_register_in(pmt::pmt_t port) { d_in_ports[pmt::symbol_to_string(port)]; }
    void message_port_register_out(pmt::pmt_t port) { d_out_ports[pmt::symbol_to_string(port)]; }
    template <typename F> void set_msg_handler(pmt::pmt_t port, F f) { d_in_ports[pmt::symbol_to_string(port)] = msg_handler_t(f); }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg) {
        std::map<std::string, std::vector<std::pair<basic_block *, std::string>>>::iterator it = d_out_ports.find(pmt::symbol_to_string(port));
        if (it == d_out_ports.end()) return;
        for (size_t i = 0; i < it->second.size(); ++i) it->second[i].first->dispatch_msg(it->second[i].second, msg);
    }
    bool has_msg_port_in(const std::string &p) const { return d_in_ports.count(p) != 0; }
    bool has_msg_port_out(const std::string &p) const { return d_out_ports.count(p) != 0; }
    void dispatch_msg(const std::string &port, pmt::pmt_t msg) {
        std::map<std::string, msg_handler_t>::iterator it = d_in_ports.find(port);
        if (it != d_in_ports.end() && it->second) it->second(msg);
    }
    // flowgraph wiring (top_block.msg_connect)
    void subscribe(const std::string &out_port, basic_block *dst, const std::string &in_port) { d_out_ports[out_port].push_back(std::make_pair(dst, in_port)); }
private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    std::map<std::string, msg_handler_t> d_in_ports;
    std::map<std::string, std::vector<std::pair<basic_block *, std::string>>> d_out_ports;
};

class block : public basic_block {
public:
    enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
    block() {}
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : basic_block(name, in, out) {}
    void consume_each(int n) { d_consumed += n; }
    long nitems_consumed() const { return d_consumed; }
    // scheduler hints (recorded; the QA driver honours them the way the GNU Radio scheduler would)
    void set_max_noutput_items(int m) { d_max_noutput = m; }
    int max_noutput_items() const { return d_max_noutput; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    virtual void forecast(int, gr_vector_int &) {}
    virtual int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &) { return noutput_items; }
private:
    long d_consumed = 0;
    int d_max_noutput = 0, d_output_multiple = 1;
};

class sync_block : public block {
public:
    sync_block() {}
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
};

inline void msg_connect(basic_block &src, const std::string &out_port, basic_block &dst, const std::string &in_port) {
    src.subscribe(out_port, &dst, in_port);
}

}  // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }     // GNU Radio 3.7: boost::shared_ptr
}
