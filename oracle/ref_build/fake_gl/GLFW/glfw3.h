/* TEST INFRASTRUCTURE: just enough of <GLFW/glfw3.h> for the reference's headers to compile headless. */
#ifndef FAKE_GLFW3_H
#define FAKE_GLFW3_H
typedef struct GLFWwindow GLFWwindow;
#endif
