#pragma once
namespace message_filters { namespace sync_policies {
template <typename M0, typename M1, typename M2> struct ApproximateTime {
    explicit ApproximateTime(uint32_t) {}
    typedef M0 Msg0; typedef M1 Msg1; typedef M2 Msg2;
};
}}  // namespace message_filters::sync_policies
