/* TEST INFRASTRUCTURE (oracle/_ref build only): stands in for a third-party header that is not vendored under /root/reference. */
