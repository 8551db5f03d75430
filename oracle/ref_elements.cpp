// ref_elements.cpp -- placeholder, filled in below
