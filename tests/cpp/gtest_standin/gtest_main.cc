// main() of a test program built against the GoogleTest stand-in (gtest/gtest.h next to this file)
#include <gtest/gtest.h>

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
