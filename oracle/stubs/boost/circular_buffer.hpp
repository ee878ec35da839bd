/* Oracle stub for <boost/circular_buffer.hpp>: the five members the reference's
 * sliding window uses (test infrastructure, not product code). */
#ifndef ORACLE_STUB_BOOST_CIRCULAR_BUFFER_HPP
#define ORACLE_STUB_BOOST_CIRCULAR_BUFFER_HPP
#include <cstddef>
#include <vector>
namespace boost {
template <class T> class circular_buffer
{
public:
    circular_buffer() : d_head(0), d_size(0) {}
    void set_capacity(size_t c)
    {
        d_buf.assign(c, T());
        d_head = 0;
        d_size = 0;
    }
    size_t size() const { return d_size; }
    size_t capacity() const { return d_buf.size(); }
    void push_back(const T &v)
    {
        if (d_buf.empty()) return;
        if (d_size == d_buf.size()) { /* full: overwrite the oldest element */
            d_buf[d_head] = v;
            d_head = (d_head + 1) % d_buf.size();
        } else {
            d_buf[(d_head + d_size) % d_buf.size()] = v;
            d_size++;
        }
    }
    T &front() { return d_buf[d_head]; }
    void pop_front()
    {
        d_head = (d_head + 1) % d_buf.size();
        d_size--;
    }
    T &operator[](size_t i) { return d_buf[(d_head + i) % d_buf.size()]; }

private:
    std::vector<T> d_buf;
    size_t d_head, d_size;
};
} // namespace boost
#endif
