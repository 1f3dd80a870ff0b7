// ORACLE — TEST INFRASTRUCTURE ONLY.  framegrabber/framegrabber.h declares a boost::mutex member; the
// grabber itself is not compiled into the oracle.
#ifndef SL2_ORACLE_BOOST_THREAD_STUB
#define SL2_ORACLE_BOOST_THREAD_STUB
namespace boost {
struct mutex {
  struct scoped_lock {
    explicit scoped_lock(mutex &) {}
  };
};
struct thread {};
}  // namespace boost
#endif
